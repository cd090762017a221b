#!/bin/bash
mkdir -p gpurun_out
run() {
  name=$1; shift
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/bench_$name.json 2>gpurun_out/bench_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$name.json")); print("$name value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), {k[:6]:round(v["ms_per_frame"],3) for k,v in d["kernel_breakdown"].items() if k[:4] in ("gemm","laye")})
except Exception as e: print("$name ERR", e); print(open("gpurun_out/bench_$name.err").read()[-2000:])
PY
}
run base
SMB_PLAN_DIV=4 SMB_FC1_SWAP=0 run div4
SMB_PLAN_DIV=4 SMB_FC1_SWAP=1 run div4_fc1swap
SMB_PLAN_DIV=4 SMB_FC1_SWAP=0 SMB_SPLIT_SMS=80 run div4_split80
SMB_PLAN_DIV=2 SMB_FC1_SWAP=0 SMB_SPLIT_SMS=80 run div2_split80
SMB_PLAN_DIV=4 SMB_FC1_SWAP=0 SMB_SPLITK=1 run div4_nosplit
