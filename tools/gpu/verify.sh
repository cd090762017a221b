#!/bin/bash
# smoke() + the whole -m gpu suite (one process) + the default bench; logs in gpurun_out/
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 -p no:cacheprovider > gpurun_out/pytest_gpu_all.log 2>&1; echo "pytest exit $?"
tail -4 gpurun_out/pytest_gpu_all.log
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit $?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_default.json"))
print({k:d[k] for k in ("value","ms_per_step","tokens_per_s","gpu_launches")}, "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "ms/token", d["roofline"]["avg_launch_ms"], "clocks", d["clocks"])
PY
tail -3 gpurun_out/bench_default.err
