#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --workload dense_decode --steps 512 --warmup 3 --no-cpu-baseline --no-frames-stage > gpurun_out/bench_dense_r02.json 2> gpurun_out/bench_dense_r02.err; echo "dense exit $?"
tail -c 1800 gpurun_out/bench_dense_r02.json; tail -3 gpurun_out/bench_dense_r02.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_kv -s 64 -c 4 -o gpurun_out/prof_attn_kv_r02 -f \
    python tools/prefill_probe.py --ctx 8000 --new 11 --reps 1 > gpurun_out/ncu_attn_kv_r02.log 2>&1
tail -2 gpurun_out/ncu_attn_kv_r02.log; ls -la gpurun_out/prof_attn_kv_r02.ncu-rep
