#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/preprocess_bench.py 2>&1 | tail -4
timeout 200 ncu --set full --clock-control none --import-source on -k regex:preprocess -s 8 -c 2 -f -o gpurun_out/prof_preprocess python tools/preprocess_bench.py > gpurun_out/ncu_pre.log 2>&1
ls -la gpurun_out/prof_preprocess.ncu-rep
