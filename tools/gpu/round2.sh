#!/bin/bash
# round-2 evidence in one GPU call: reference arm, dense workload, ncu launch list of the default bench command (bounded), --set full captures
# of the decode kernel and of the prefill attention kernel (fire-sized call); outputs in gpurun_out/ (summarised into profiles/ by hand)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r02_ref.json 2> gpurun_out/bench_r02_ref.err
tail -c 400 gpurun_out/bench_r02_ref.json
timeout 900 python bench.py --workload dense_decode --steps 512 --warmup 3 --no-cpu-baseline --no-frames-stage > gpurun_out/bench_dense_r02.json 2> gpurun_out/bench_dense_r02.err; echo "dense exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r02.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graphs --no-frames-stage > gpurun_out/ncu_launch_r02.log 2>&1
wc -l gpurun_out/launches_r02.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_stream -s 20 -c 2 -o gpurun_out/prof_decode_r02 -f \
    python tools/decode_probe.py --layers 32 --ctx 2048 > gpurun_out/ncu_decode_r02.log 2>&1
tail -2 gpurun_out/ncu_decode_r02.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_kv -s 576 -c 4 -o gpurun_out/prof_attn_kv_fire_r02 -f \
    python tools/prefill_probe.py --ctx 8000 --new 11 --reps 1 > gpurun_out/ncu_attn_kv_fire_r02.log 2>&1
tail -2 gpurun_out/ncu_attn_kv_fire_r02.log
ls -la gpurun_out | tail -8
