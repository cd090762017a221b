#!/bin/bash
# decode-kernel probe: product build and the wait-probe build at ctx 2048; output gpurun_out/ds_sweep.txt
mkdir -p gpurun_out
out=gpurun_out/ds_sweep.txt; : > $out
echo "== product build ctx 2048" >> $out
timeout 300 python tools/decode_probe.py --layers 32 --ctx 2048 --phases 2>&1 | grep -v Warning | tail -10 >> $out
echo "== wait-probe build ctx 2048" >> $out
SMB_LIB_PATH=$PWD/streammind_b200/libstreammind_b200_probe.so timeout 300 python tools/decode_probe.py --layers 32 --ctx 2048 --phases 2>&1 | grep -v Warning | tail -10 >> $out
echo "== wait-probe build ctx 2048, math skipped" >> $out
SMB_DS_DBG=1 SMB_LIB_PATH=$PWD/streammind_b200/libstreammind_b200_probe.so timeout 300 python tools/decode_probe.py --layers 32 --ctx 2048 --phases 2>&1 | grep -v Warning | tail -10 >> $out
echo "== prefill probe" >> $out
timeout 300 python tools/prefill_probe.py --ctx 2048 8000 --new 11 30 2>&1 | grep -v Warning | tail -6 >> $out
cat $out
