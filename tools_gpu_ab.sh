#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_frame_path_gpu.py tests/test_stream_api_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -6
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_a.json 2>gpurun_out/bench_a.err; tail -3 gpurun_out/bench_a.err
SMB_SPLITK=1 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b.json 2>gpurun_out/bench_b.err
python bench.py --steps 3 --warmup 3 --chunk 8 --no-cpu-baseline > gpurun_out/bench_c8.json 2>gpurun_out/bench_c8.err
python - <<'PY'
import json
for f in ("bench_a","bench_b","bench_c8"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, "value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), {k:round(v["ms_per_frame"],3) for k,v in d["kernel_breakdown"].items()}, "gemm frac", round(d["roofline"]["frac"],3))
    except Exception as e: print(f, "ERR", e)
PY
