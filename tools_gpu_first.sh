#!/bin/bash
# first GPU contact: each test file in its own process so a trap in one does not poison the next
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for f in test_gemm_gpu test_attention_gpu test_frame_path_gpu; do
  timeout 900 python -m pytest tests/$f.py -m gpu -q -x --timeout 600 -p no:cacheprovider > gpurun_out/$f.log 2>&1
  echo "$f exit $?" >> gpurun_out/summary.txt
  tail -5 gpurun_out/$f.log
done
cat gpurun_out/summary.txt
