#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
echo "== ctx 2048"
timeout 120 python tools/decode_probe.py --layers 32 --ctx 2048 --phases 2>&1 | tail -6
echo "== ctx 4096"
timeout 120 python tools/decode_probe.py --layers 32 --ctx 4096 --phases 2>&1 | tail -1
echo "== 2 streams"
timeout 120 python tools/decode_probe.py --layers 32 --ctx 2048 --streams 2 --phases 2>&1 | tail -1
echo "== 4 streams"
timeout 120 python tools/decode_probe.py --layers 32 --ctx 2048 --streams 4 --phases 2>&1 | tail -1
