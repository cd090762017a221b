#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_frame_path_gpu.py tests/test_stream_api_gpu.py -m gpu -q -x --timeout 900 -p no:cacheprovider > gpurun_out/pytest_lanes4.log 2>&1
echo "pytest exit $?"; tail -8 gpurun_out/pytest_lanes4.log
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
run l4_div4 --lookahead 4
SMB_LANES=2 run l2_div4
SMB_LANES=3 run l3_div4 --lookahead 3
SMB_PLAN_DIV=8 run l4_div8 --lookahead 4
SMB_PLAN_DIV=8 SMB_SPLIT_SMS=80 run l4_div8_s80 --lookahead 4
SMB_PLAN_DIV=8 SMB_SPLIT_SMS=40 run l4_div8_s40 --lookahead 4
SMB_PLAN_DIV=2 run l4_div2 --lookahead 4
