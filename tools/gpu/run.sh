#!/bin/bash
# usage: tools_gpu_run.sh <test files...>; each in its own process
mkdir -p gpurun_out
for f in "$@"; do
  timeout 1200 python -m pytest tests/$f.py -m gpu -q -x --timeout 900 -p no:cacheprovider -s > gpurun_out/$f.log 2>&1
  echo "$f exit $?"
  tail -15 gpurun_out/$f.log
done
