#!/bin/bash
mkdir -p gpurun_out
run() {
  name=$1; shift
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/bench_$name.json 2>gpurun_out/bench_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$name.json")); print("$name value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "serial", round(d["serial_b1"]["value"],1), "attn ms/frame", round(d["kernel_breakdown"]["attention_kernel"]["ms_per_frame"],3))
except Exception as e: print("$name ERR", e); print(open("gpurun_out/bench_$name.err").read()[-2000:])
PY
}
run base
SMB_LIB_PATH=$PWD/streammind_b200/libsmb_a3.so run attn3
SMB_LIB_PATH=$PWD/streammind_b200/libsmb_a3.so timeout 600 python -m pytest tests/test_attention_gpu.py tests/test_frame_path_gpu.py -m gpu -q -x -p no:cacheprovider -k "attention or full_width" 2>&1 | tail -2
