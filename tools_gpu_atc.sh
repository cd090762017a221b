#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_attention_gpu.py -m gpu -x -q 2>&1 | tail -3
rc=${PIPESTATUS[0]}
if [ $rc -ne 0 ]; then echo "ATTENTION TEST FAILED rc=$rc"; exit 0; fi
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/nopoly.json 2> gpurun_out/nopoly.err
python - <<PY
import json
d=json.loads(open("gpurun_out/nopoly.json").read().strip().splitlines()[-1])
print("no poly, no stamps:", d["value"], d["e2e"]["value"], d["kernel_breakdown"]["attention_kernel"]["ms_per_frame"])
PY
