#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke.log
timeout 400 python bench.py > gpurun_out/bench_tc_attn.json 2> gpurun_out/bench_tc_attn.err; tail -c 400 gpurun_out/bench_tc_attn.json
timeout 300 python tools/attn_trace.py 2>&1 | tail -13
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_tc_attn.csv \
    python bench.py --steps 1 --warmup 3 --no-graphs --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 30 -c 1 -f -o gpurun_out/atc_full2 \
    python bench.py --steps 1 --warmup 3 --no-graphs --no-cpu-baseline > gpurun_out/ncu_atc.log 2>&1
ls -la gpurun_out | tail -8
