#!/bin/bash
# end-of-round validation: smoke, decode workloads (profiles), default bench
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -6 gpurun_out/smoke.log
timeout 900 python bench.py --workload gated_decode --steps 1 --warmup 3 > gpurun_out/bench_gated_decode.json 2>gpurun_out/bench_gated_decode.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_gated_decode.json")); print("gated_decode fps", round(d["value"],2), d["decode"]["tokens_per_s"], d["roofline"]["frac"], d["stream_roofline"]["frac"])
PY
timeout 900 python bench.py --workload dense_decode --steps 1 --warmup 3 > gpurun_out/bench_dense_decode.json 2>gpurun_out/bench_dense_decode.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_dense_decode.json")); print("dense_decode fps", round(d["value"],2), d["decode"]["tokens_per_s"], d["roofline"]["frac"], d["stream_roofline"]["frac"])
PY
