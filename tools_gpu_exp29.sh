#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 -p no:cacheprovider > gpurun_out/pytest_final.log 2>&1
echo "pytest exit $?"; tail -12 gpurun_out/pytest_final.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_q.json 2>gpurun_out/bench_q.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_q.json")); print("value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "serial", round(d["serial_b1"]["value"],1))
PY
