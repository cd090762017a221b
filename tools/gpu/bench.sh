#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
tail -c 3000 gpurun_out/bench_default.json; tail -5 gpurun_out/bench_default.err
python bench.py --steps 3 --warmup 3 --chunk 8 --no-cpu-baseline > gpurun_out/bench_chunk8.json 2> gpurun_out/bench_chunk8.err
python bench.py --steps 3 --warmup 3 --no-graphs --no-cpu-baseline > gpurun_out/bench_nographs.json 2> gpurun_out/bench_nographs.err
# launch list (cold-cache, serialised): 1 frame after 2 warm frames
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --frames 1 --no-graphs --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
# full capture of one gemv (gate) and one gemm launch
ncu --set full --clock-control none --import-source on -k regex:gemv_kernel -s 40 -c 4 -o gpurun_out/prof_gemv -f \
    python bench.py --steps 1 --warmup 3 --frames 1 --no-graphs --no-cpu-baseline > gpurun_out/ncu_gemv.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 100 -c 4 -o gpurun_out/prof_gemm -f \
    python bench.py --steps 1 --warmup 3 --frames 1 --no-graphs --no-cpu-baseline > gpurun_out/ncu_gemm.log 2>&1
ls -la gpurun_out
