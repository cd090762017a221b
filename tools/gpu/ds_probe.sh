#!/bin/bash
# decode-kernel probes on the GPU box: the product build (per-CTA phase clocks), the exchange chain alone, the consumers alone (chunks fed from L2),
# and -- when tools/gpu/ds_probe_build.sh was run before the call -- the wait / attention sub-phase probes.  Output: gpurun_out/ds_probe.txt
mkdir -p gpurun_out
out=gpurun_out/ds_probe.txt; : > $out
for args in "--ctx 2048 --phases" "--ctx 4096" "--ctx 8000" "--ctx 2048 --streams 2" "--ctx 2048 --streams 4"; do
  timeout 300 python tools/decode_probe.py --layers 32 $args 2>&1 | grep -E "^layers|CTA [01] " >> $out
done
for f in 2 16 17; do
  echo "== SMB_DS_DBG=$f (2: no weight stream, 16: weight chunks from L2, 17: 16 + math skipped)" >> $out
  SMB_DS_DBG=$f timeout 300 python tools/decode_probe.py --layers 32 --ctx 2048 --phases 2>&1 | grep -E "^layers|CTA [01] " >> $out
done
if [ -f streammind_b200/libstreammind_b200_probe.so ]; then
  echo "== probe build (-DSMB_DS_WAITPROBE)" >> $out
  SMB_LIB_PATH=$PWD/streammind_b200/libstreammind_b200_probe.so timeout 300 python tools/decode_probe.py --layers 32 --ctx 2048 --phases 2>&1 | grep -E "^layers|CTA [01] |per-CTA" >> $out
fi
cat $out
