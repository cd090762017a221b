#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x --timeout 300 -p no:cacheprovider > gpurun_out/pytest_bm2.log 2>&1
echo "pytest gemm exit $?"; tail -8 gpurun_out/pytest_bm2.log
timeout 900 python -m pytest tests/test_frame_path_gpu.py tests/test_stream_api_gpu.py tests/test_llm_gpu.py -m gpu -q -x --timeout 600 -p no:cacheprovider > gpurun_out/pytest_bm2b.log 2>&1
echo "pytest frame exit $?"; tail -8 gpurun_out/pytest_bm2b.log
run() {
  name=$1; shift
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/bench_$name.json 2>gpurun_out/bench_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$name.json")); print("$name value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1))
except Exception as e: print("$name ERR", e); print(open("gpurun_out/bench_$name.err").read()[-2000:])
PY
}
run bm2
SMB_BM2=0 run nobm2
run bm2_chunk8 --chunk 8
SMB_BM2=0 run nobm2_chunk8 --chunk 8
run bm2_serial --no-pipeline
