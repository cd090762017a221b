#!/bin/bash
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 40 -c 5 -o gpurun_out/prof_gemm_chunk8 -f \
    python bench.py --steps 1 --warmup 3 --frames 8 --chunk 8 --no-pipeline --no-graphs --no-cpu-baseline > gpurun_out/ncu_gemm_c8.log 2>&1
ls -la gpurun_out/prof_gemm_chunk8.ncu-rep
