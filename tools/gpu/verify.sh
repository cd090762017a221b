#!/bin/bash
# full -m gpu suite (one process) + the default bench; logs in gpurun_out/
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 -p no:cacheprovider > gpurun_out/pytest_gpu_all.log 2>&1; echo "pytest exit $?"
tail -6 gpurun_out/pytest_gpu_all.log
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit $?"
tail -c 2500 gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err
