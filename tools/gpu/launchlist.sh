#!/bin/bash
# ncu launch list of the default bench command, this library's kernels only (warm-up step + timed step, bounded)
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_r02.csv \
    -k regex:"attention|decode_stream|gemm_tc|gemv|layernorm|rmsnorm|swiglu|rope|splitk|gather|argmax|mamba|vit_|im2col|preprocess|gqa|ds_first|cos_sim|topk" \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graphs --no-frames-stage > gpurun_out/ncu_launch_r02.log 2>&1
wc -l gpurun_out/launches_r02.csv; tail -c 300 gpurun_out/ncu_launch_r02.log
