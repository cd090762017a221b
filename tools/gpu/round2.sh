#!/bin/bash
# round-2 evidence in one GPU call: default bench (configs[2]), reference arm, ncu launch list of the same command (bounded),
# one --set full capture of the decode kernel; outputs in gpurun_out/ (copied to profiles/ by hand)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python bench.py --steps 16 --warmup 3 --with-frames-stage > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err
tail -c 1500 gpurun_out/bench_r02.json; tail -3 gpurun_out/bench_r02.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r02_ref.json 2> gpurun_out/bench_r02_ref.err
tail -c 600 gpurun_out/bench_r02_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r02.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graphs > gpurun_out/ncu_launch_r02.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_stream -s 20 -c 2 -o gpurun_out/prof_decode_r02 -f \
    python tools/decode_probe.py --layers 32 --ctx 2048 > gpurun_out/ncu_decode_r02.log 2>&1
tail -2 gpurun_out/ncu_decode_r02.log
ls -la gpurun_out | tail -8
