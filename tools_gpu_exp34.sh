#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_frame_path_gpu.py -m gpu -q -x --timeout 900 -p no:cacheprovider -s -k "batched_gate_gemm" > gpurun_out/pytest_gg2.log 2>&1
echo "pytest exit $?"; grep -E "gate vs oracle|gate-as-GEMM|passed|failed|Error|assert" gpurun_out/pytest_gg2.log | tail -12
