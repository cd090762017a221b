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
SMB_SPLITK=2 SMB_GEMM_PRE=0 run s2_nopre
SMB_SPLITK=1 run s1
SMB_SPLITK=2 SMB_LANES=6 run s2_l6 --lookahead 7
SMB_SPLITK=2 SMB_LANES=8 run s2_l8 --lookahead 7
SMB_SPLITK=2 SMB_PLAN_DIV=8 run s2_div8
SMB_SPLITK=2 SMB_FC1_SWAP=0 run s2_b
