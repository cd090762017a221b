#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/ds_sweep.txt; : > $out
timeout 300 python tools/decode_probe.py --layers 32 --ctx 2048 --phases 2>&1 | grep -E "^layers|CTA [01] " >> $out
export SMB_LIB_PATH=$PWD/streammind_b200/libstreammind_b200_probe.so
SMB_DS_DBG=2 timeout 300 python tools/decode_probe.py --layers 32 --ctx 2048 --phases 2>&1 | grep -E "^layers|CTA [01] " >> $out
timeout 300 python tools/decode_probe.py --layers 32 --ctx 2048 --phases 2>&1 | grep -E "^layers|CTA [01] " >> $out
cat $out
