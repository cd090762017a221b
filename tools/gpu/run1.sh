#!/bin/bash
# usage: tools/gpu/run1.sh <pytest args...>
mkdir -p gpurun_out
timeout 1500 python -m pytest "$@" -m gpu -q --timeout 600 -p no:cacheprovider > gpurun_out/pytest_run1.log 2>&1; echo "pytest exit $?"
tail -30 gpurun_out/pytest_run1.log
