#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --workload gated_decode --steps 1 --warmup 3 > gpurun_out/bench_gated_decode.json 2>gpurun_out/bench_gated_decode.err
tail -c 1800 gpurun_out/bench_gated_decode.json; tail -3 gpurun_out/bench_gated_decode.err
timeout 900 python bench.py --workload dense_decode --steps 1 --warmup 3 > gpurun_out/bench_dense_decode.json 2>gpurun_out/bench_dense_decode.err
tail -c 1800 gpurun_out/bench_dense_decode.json; tail -3 gpurun_out/bench_dense_decode.err
