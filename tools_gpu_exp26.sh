#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemv_seg_kernel -s 14 -c 4 -o gpurun_out/prof_gemv_seg -f \
    python bench.py --steps 1 --warmup 3 --frames 8 --no-graphs --no-cpu-baseline > gpurun_out/ncu_gemv_seg.log 2>&1
ls -la gpurun_out/prof_gemv_seg.ncu-rep
