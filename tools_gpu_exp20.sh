#!/bin/bash
mkdir -p gpurun_out
run() {
  name=$1; shift
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/bench_$name.json 2>gpurun_out/bench_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$name.json")); print("$name value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1))
except Exception as e: print("$name ERR", e); print(open("gpurun_out/bench_$name.err").read()[-2000:])
PY
}
run base
SMB_PLAN_DIV=2 run div2
SMB_PLAN_DIV=8 run div8
SMB_SPLITK=2 run split2
SMB_SPLITK=3 run split3
SMB_LANES=3 run lanes3
SMB_LANES=2 run lanes2
SMB_NO_PDL=1 run nopdl
SMB_GEMM_PRE=0 run nopre
