#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_llm_gpu.py tests/test_llm_fullsize_gpu.py -m gpu -x -q 2>&1 | tail -3
echo "== ctx 2048"
timeout 120 python tools/decode_probe.py --layers 32 --ctx 2048 --phases 2>&1 | tail -6
echo "== nomath"
SMB_DS_DBG=1 timeout 120 python tools/decode_probe.py --layers 32 --ctx 2048 --phases 2>&1 | tail -1
echo "== 2 streams"
timeout 120 python tools/decode_probe.py --layers 32 --ctx 2048 --streams 2 --phases 2>&1 | tail -1
