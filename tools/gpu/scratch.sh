#!/bin/bash
timeout 200 python -m pytest tests/test_preprocess_gpu.py -m gpu -x -q 2>&1 | tail -3
timeout 100 python tools/preprocess_bench.py 2>&1 | tail -3
timeout 300 python -m pytest tests -m gpu -x -q --deselect tests/test_preprocess_gpu.py 2>&1 | tail -2
