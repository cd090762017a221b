#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_llm_gpu.py tests/test_stream_api_gpu.py -m gpu -q -x --timeout 600 -p no:cacheprovider > gpurun_out/pytest_pf.log 2>&1
echo "pytest exit $?"; tail -5 gpurun_out/pytest_pf.log
timeout 900 python bench.py --workload dense_decode --frames 128 --steps 1 --warmup 3 > gpurun_out/bench_dd_a.json 2>gpurun_out/bench_dd_a.err
SMB_PREFILL_SPLITK=0 timeout 900 python bench.py --workload dense_decode --frames 128 --steps 1 --warmup 3 > gpurun_out/bench_dd_b.json 2>gpurun_out/bench_dd_b.err
python - <<PY
import json
for n in ("a","b"):
    d=json.load(open(f"gpurun_out/bench_dd_{n}.json")); print("dense_decode 128f splitk", n, "fps", round(d["value"],2), d["decode"]["tokens_per_s"])
PY
