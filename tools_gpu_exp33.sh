#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_frame_path_gpu.py tests/test_stream_api_gpu.py -m gpu -q -x --timeout 900 -p no:cacheprovider -s > gpurun_out/pytest_gg.log 2>&1
echo "pytest exit $?"; grep -E "gate-as-GEMM|tower-batch|interleaved|passed|failed|Error" gpurun_out/pytest_gg.log | tail -16
run() {
  name=$1; shift
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/bench_$name.json 2>gpurun_out/bench_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$name.json")); print("$name value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), {k[:6]:round(v["ms_per_frame"],3) for k,v in d["kernel_breakdown"].items() if k[:4] in ("gemm","gemv")})
except Exception as e: print("$name ERR", e); print(open("gpurun_out/bench_$name.err").read()[-2000:])
PY
}
run gg
SMB_GATE_GEMM=0 run nogg
run gg_chunk8 --chunk 8
