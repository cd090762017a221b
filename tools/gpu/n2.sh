#!/bin/bash
# the driver's N = 2 launch of both arms
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "n2 exit $?"
tail -c 1200 gpurun_out/bench_n2.json; tail -3 gpurun_out/bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_n2_ref.json 2> gpurun_out/bench_n2_ref.err; echo "n2 ref exit $?"
tail -c 300 gpurun_out/bench_n2_ref.json
