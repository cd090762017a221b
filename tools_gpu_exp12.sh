#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 900 -p no:cacheprovider > gpurun_out/pytest_pre.log 2>&1
echo "pytest exit $?"; tail -8 gpurun_out/pytest_pre.log
run() {
  name=$1; shift
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/bench_$name.json 2>gpurun_out/bench_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$name.json")); print("$name value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), {k[:6]:round(v["ms_per_frame"],3) for k,v in d["kernel_breakdown"].items()})
except Exception as e: print("$name ERR", e); print(open("gpurun_out/bench_$name.err").read()[-2000:])
PY
}
run serial --no-pipeline
run pipe
SMB_GEMM_PRE=0 run serial_nopre --no-pipeline
SMB_GEMM_PRE=0 run pipe_nopre
run pipe_chunk8 --chunk 8
