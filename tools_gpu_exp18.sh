#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 900 -p no:cacheprovider > gpurun_out/pytest_nv.log 2>&1
echo "pytest exit $?"; tail -12 gpurun_out/pytest_nv.log
run() {
  name=$1; shift
  timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/bench_$name.json 2>gpurun_out/bench_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$name.json")); print("$name value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "gemv ms/frame", round(d["kernel_breakdown"]["gemv_kernel"]["ms_per_frame"],3))
except Exception as e: print("$name ERR", e); print(open("gpurun_out/bench_$name.err").read()[-2000:])
PY
}
run pipe_b1
run pipe_chunk4 --chunk 4
run pipe_chunk8 --chunk 8
run serial_chunk8 --chunk 8 --no-pipeline
SMB_GEMV_BATCH=2 run pipe_chunk8_nv2 --chunk 8
SMB_GEMV_BATCH=1 run pipe_chunk8_nv1 --chunk 8
