#!/bin/bash
mkdir -p gpurun_out
run() {
  name=$1; shift
  timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/bench_$name.json 2>gpurun_out/bench_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$name.json")); print("$name value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1))
except Exception as e: print("$name ERR", e); print(open("gpurun_out/bench_$name.err").read()[-2000:])
PY
}
SMB_BM2=0 SMB_MEGA=2 run mega_chunk8_serial --chunk 8 --no-pipeline
SMB_BM2=0 SMB_MEGA=0 run nomega_chunk8_serial --chunk 8 --no-pipeline
SMB_BM2=0 SMB_MEGA=2 SMB_MEGA_BN=256 run mega256_chunk8_serial --chunk 8 --no-pipeline
SMB_BM2=0 SMB_MEGA=2 SMB_LANES=1 run mega_chunk8_pipe1 --chunk 8
SMB_BM2=0 SMB_MEGA=2 run mega_chunk16_serial --chunk 16 --no-pipeline
SMB_BM2=0 SMB_MEGA=0 run nomega_chunk16_serial --chunk 16 --no-pipeline
