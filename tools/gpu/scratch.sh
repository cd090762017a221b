#!/bin/bash
for cfg in "0 100000" "1 100000" "2 100000" "3 100000" "0 1000" "0 100" "1 100"; do
set -- $cfg
echo "== wait $1 hint $2"
SMB_DS_WAIT=$1 SMB_DS_HINT_NS=$2 timeout 120 python tools/decode_probe.py --layers 32 --ctx 2048 --phases 2>&1 | tail -1
done
echo "== nomath wait 1"
SMB_DS_DBG=1 SMB_DS_WAIT=1 timeout 120 python tools/decode_probe.py --layers 32 --ctx 2048 --phases 2>&1 | tail -1
echo "== nomath wait 3"
SMB_DS_DBG=1 SMB_DS_WAIT=3 timeout 120 python tools/decode_probe.py --layers 32 --ctx 2048 --phases 2>&1 | tail -1
