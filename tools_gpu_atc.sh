#!/bin/bash
timeout 300 python -m pytest tests/test_preprocess_gpu.py -m gpu -x -q -s 2>&1 | tail -12
