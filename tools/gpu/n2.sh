#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "exit $?"; wc -l gpurun_out/bench_n2.json; tail -c 600 gpurun_out/bench_n2.json | head -c 600; tail -3 gpurun_out/bench_n2.err
