#!/bin/bash
# the driver's N = 8 launch of our arm (short)
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 4 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "n8 exit $?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_n8.json")); print(d["value"], d["n_gpus"], d["ms_per_step"], d["e2e"]["value"], d["tokens_per_s"], d["scaling"])
PY
tail -3 gpurun_out/bench_n8.err
