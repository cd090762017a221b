#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_r02.csv \
    -k regex:"attention|decode_stream|ds_first|gather_rows|gemm_tc|gemv|gqa_expand|im2col|layernorm|mamba|rmsnorm_rows|rope_append|splitk|swiglu|vit_|preprocess" \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graphs --no-frames-stage > gpurun_out/ncu_launch_r02.log 2>&1
tail -c 300 gpurun_out/ncu_launch_r02.log; wc -l gpurun_out/launches_r02.csv
