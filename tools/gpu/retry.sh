#!/bin/bash
# usage: tools/gpu/retry.sh <timeout_s> <command...>   -- re-submits a gpurun call while the pod answers busy (exit 3)
t=$1; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout $t -- "$@" > /tmp/gpurun_retry.log 2>&1; rc=$?
  if ! grep -q "status=transient" /tmp/gpurun_retry.log; then break; fi
  sleep 60
done
tail -${TAIL:-80} /tmp/gpurun_retry.log
exit $rc
