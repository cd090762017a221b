#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
tail -c 6000 gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err
timeout 300 python bench.py --steps 3 --warmup 3 --chunk 8 --no-cpu-baseline > gpurun_out/bench_chunk8.json 2> gpurun_out/bench_chunk8.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --frames 8 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 400 -c 8 -o gpurun_out/prof_gemm_pipe -f \
    python bench.py --steps 1 --warmup 3 --frames 8 --no-graphs --no-cpu-baseline > gpurun_out/ncu_gemm.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemv_kernel -s 60 -c 6 -o gpurun_out/prof_gemv_batched -f \
    python bench.py --steps 1 --warmup 3 --frames 8 --no-graphs --no-cpu-baseline > gpurun_out/ncu_gemv.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_kernel -s 60 -c 2 -o gpurun_out/prof_attn2 -f \
    python bench.py --steps 1 --warmup 3 --frames 8 --no-graphs --no-cpu-baseline > gpurun_out/ncu_attn.log 2>&1
ls -la gpurun_out/*.ncu-rep
