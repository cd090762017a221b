#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kv_attention_gpu.py -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/pytest_kvattn.log 2>&1; echo "kvattn exit $?"
tail -15 gpurun_out/pytest_kvattn.log
