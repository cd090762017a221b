#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_frame_path_gpu.py tests/test_stream_api_gpu.py -m gpu -q -x --timeout 900 -p no:cacheprovider > gpurun_out/pytest_tb.log 2>&1
echo "pytest exit $?"; tail -12 gpurun_out/pytest_tb.log
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
run tb8
SMB_TOWER_BATCH=4 run tb4
SMB_TOWER_BATCH=1 run tb1
SMB_LANES=3 run tb8_l3
SMB_TOWER_BATCH=8 run tb8_la7 --lookahead 7
