#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 900 -p no:cacheprovider > gpurun_out/pytest_pdl.log 2>&1
echo "pytest exit $?"; tail -8 gpurun_out/pytest_pdl.log
run() {
  name=$1; shift
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/bench_$name.json 2>gpurun_out/bench_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$name.json")); print("$name value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), {k[:6]:round(v["ms_per_frame"],3) for k,v in d["kernel_breakdown"].items()})
except Exception as e: print("$name ERR", e); print(open("gpurun_out/bench_$name.err").read()[-2000:])
PY
}
run serial --no-pipeline
run pipe
SMB_NO_PDL=1 run serial_nopdl --no-pipeline
timeout 600 python bench.py --workload gated_decode --frames 32 --steps 1 --warmup 3 > gpurun_out/bench_dec_pdl.json 2>gpurun_out/bench_dec_pdl.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_dec_pdl.json")); print("decode fps", round(d["value"],2), d["decode"], d["roofline"]["frac"])
except Exception as e: print("ERR", e); print(open("gpurun_out/bench_dec_pdl.err").read()[-1500:])
PY
