#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_frame_path_gpu.py tests/test_stream_api_gpu.py -m gpu -q -x --timeout 900 -p no:cacheprovider > gpurun_out/pytest_lanes.log 2>&1
echo "pytest exit $?"; tail -8 gpurun_out/pytest_lanes.log
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
run lanes2
run lanes2_la1 --lookahead 1
run lanes2_la3 --lookahead 3
SMB_LANES=1 run lanes1
run lanes2_chunk8 --chunk 8
run lanes2_chunk2 --chunk 2
