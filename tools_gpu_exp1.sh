#!/bin/bash
mkdir -p gpurun_out
./tools/bin/mma_bench > gpurun_out/mma_bench.txt 2>&1; cat gpurun_out/mma_bench.txt
python tools/frame_gemm_trace.py > gpurun_out/frame_gemm_trace.txt 2>&1; tail -25 gpurun_out/frame_gemm_trace.txt
python tools/frame_timeline.py > gpurun_out/frame_timeline.txt 2>&1; tail -12 gpurun_out/frame_timeline.txt
for co in 0 100 ; do
  if [ $co = 0 ]; then unset SMB_CARVEOUT; else export SMB_CARVEOUT=$co; fi
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_co$co.json 2>gpurun_out/bench_co$co.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_co$co.json")); print("carveout $co value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), {k:round(v["ms_per_frame"],3) for k,v in d["kernel_breakdown"].items()})
PY
done
