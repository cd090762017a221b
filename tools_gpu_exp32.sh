#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_frame_path_gpu.py tests/test_llm_gpu.py -m gpu -q -x --timeout 600 -p no:cacheprovider > gpurun_out/pytest_2p.log 2>&1
echo "pytest exit $?"; tail -8 gpurun_out/pytest_2p.log
run() {
  name=$1; shift
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/bench_$name.json 2>gpurun_out/bench_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$name.json")); print("$name value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "serial", round(d["serial_b1"]["value"],1), "gemm ms/frame", round(d["kernel_breakdown"]["gemm_tc_kernel"]["ms_per_frame"],3))
except Exception as e: print("$name ERR", e); print(open("gpurun_out/bench_$name.err").read()[-2000:])
PY
}
run p2
SMB_GEMM_2PROD=0 run p1
run p2_chunk8 --chunk 8
SMB_GEMM_2PROD=0 run p1_chunk8 --chunk 8
