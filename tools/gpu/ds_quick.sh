#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/ds_sweep.txt; : > $out
for v in _acc4 _acc8; do
  export SMB_LIB_PATH=$PWD/streammind_b200/libstreammind_b200$v.so
  echo "== lib$v: L2-fed chunks (consumer speed), then the real step, then 4 streams" >> $out
  SMB_DS_DBG=16 timeout 300 python tools/decode_probe.py --layers 32 --ctx 2048 --phases 2>&1 | grep -E "^layers|CTA 0 " >> $out
  timeout 300 python tools/decode_probe.py --layers 32 --ctx 2048 --phases 2>&1 | grep -E "^layers|CTA 0 " >> $out
  timeout 300 python tools/decode_probe.py --layers 32 --ctx 2048 --streams 4 2>&1 | grep -E "^layers" >> $out
done
cat $out
