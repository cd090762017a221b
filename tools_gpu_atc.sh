#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_attention_gpu.py -m gpu -x -q 2>&1 | tail -6
rc=${PIPESTATUS[0]}
if [ $rc -ne 0 ]; then echo "ATTENTION TEST FAILED rc=$rc"; exit 0; fi
timeout 200 python tools/attn_trace.py 2>&1 | tail -16
timeout 600 python -m pytest tests/test_frame_path_gpu.py -m gpu -x -q -k "full_size" 2>&1 | tail -4
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/atc_on.json 2> gpurun_out/atc_on.err; tail -c 300 gpurun_out/atc_on.json
