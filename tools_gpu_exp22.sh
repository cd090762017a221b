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
SMB_PIPE_PRE=1 run pre1
SMB_PLAN_DIV=8 run div8
SMB_PLAN_DIV=16 run div16
SMB_LANES=6 run l6 --lookahead 7
SMB_LANES=8 run l8 --lookahead 7
SMB_LANES=8 SMB_PLAN_DIV=8 run l8_div8 --lookahead 7
