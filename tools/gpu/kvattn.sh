#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kv_attention_gpu.py tests/test_llm_gpu.py -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/pytest_kvattn.log 2>&1; echo "kvattn exit $?"
tail -4 gpurun_out/pytest_kvattn.log
timeout 300 python tools/prefill_probe.py --ctx 2048 8000 --new 11 30 2>&1 | grep -v Warning | tail -4
