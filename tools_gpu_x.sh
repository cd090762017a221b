#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 -p no:cacheprovider > gpurun_out/pytest_final3.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/pytest_final3.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_default.json"))
print("value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "serial", round(d["serial_b1"]["value"],1), "cpu", d["cpu_baseline"]["value"])
print("roofline", d["roofline"]["kernel"], round(d["roofline"]["achieved"],1), round(d["roofline"]["frac"],3))
print("secondary", d["roofline_secondary"]["kernel"][:40], round(d["roofline_secondary"]["achieved"],1), round(d["roofline_secondary"]["frac"],3))
print({k:round(v["ms_per_frame"],3) for k,v in d["kernel_breakdown"].items()})
PY
timeout 300 python bench.py --steps 3 --warmup 3 --chunk 8 --no-cpu-baseline > gpurun_out/bench_chunk8.json 2> gpurun_out/bench_chunk8.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_chunk8.json")); print("chunk8 value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1))
PY
