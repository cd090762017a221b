#!/bin/bash
mkdir -p gpurun_out
run() {
  name=$1; shift
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/bench_$name.json 2>gpurun_out/bench_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$name.json")); print("$name value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "gemm ms/frame", round(d["kernel_breakdown"]["gemm_tc_kernel"]["ms_per_frame"],3))
except Exception as e: print("$name ERR", e); print(open("gpurun_out/bench_$name.err").read()[-2000:])
PY
}
run base
SMB_PGEMM=3 run pg3
SMB_PGEMM=4 run pg4
SMB_PGEMM=6 run pg6
SMB_BM2=1 run bm2
SMB_PLAN_DIV=2 run div2
SMB_PIPE_PRE=1 run pre1
