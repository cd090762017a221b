#!/bin/bash
# one GPU call: parity tests, default bench, ncu launch list, ncu full captures of the three top kernels
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q -x --timeout 900 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
tail -c 4000 gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err
timeout 300 python bench.py --steps 3 --warmup 3 --chunk 8 --no-cpu-baseline > gpurun_out/bench_chunk8.json 2> gpurun_out/bench_chunk8.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --frames 1 --no-graphs --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemv_kernel -s 40 -c 6 -o gpurun_out/prof_gemv -f \
    python bench.py --steps 1 --warmup 3 --frames 1 --no-graphs --no-cpu-baseline > gpurun_out/ncu_gemv.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 100 -c 6 -o gpurun_out/prof_gemm -f \
    python bench.py --steps 1 --warmup 3 --frames 1 --no-graphs --no-cpu-baseline > gpurun_out/ncu_gemm.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_kernel -s 30 -c 2 -o gpurun_out/prof_attn -f \
    python bench.py --steps 1 --warmup 3 --frames 1 --no-graphs --no-cpu-baseline > gpurun_out/ncu_attn.log 2>&1
# pipelined path (tower chunks of 8): launch list bounded to 16 frames, then the tcgen05 attention and the preprocessing kernels
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_pipelined.csv \
    python bench.py --steps 1 --warmup 3 --frames 16 --no-graphs --no-cpu-baseline > gpurun_out/ncu_launch_pipe.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 30 -c 1 -o gpurun_out/prof_attn_tc -f \
    python bench.py --steps 1 --warmup 3 --frames 16 --no-graphs --no-cpu-baseline > gpurun_out/ncu_attn_tc.log 2>&1
timeout 120 python tools/preprocess_bench.py > gpurun_out/preprocess_bench.log 2>&1; tail -3 gpurun_out/preprocess_bench.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:preprocess -s 8 -c 2 -o gpurun_out/prof_preprocess -f \
    python tools/preprocess_bench.py > gpurun_out/ncu_pre.log 2>&1
ls -la gpurun_out
