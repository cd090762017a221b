#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_frame_path_gpu.py -m gpu -q -x --timeout 300 -p no:cacheprovider -k "pipelined" > gpurun_out/pytest_pg0.log 2>&1
echo "pytest (pgemm off) exit $?"; tail -3 gpurun_out/pytest_pg0.log
SMB_PGEMM=2 timeout 600 python -m pytest tests/test_frame_path_gpu.py -m gpu -q -x --timeout 300 -p no:cacheprovider -k "pipelined" > gpurun_out/pytest_pg2.log 2>&1
echo "pytest (pgemm 2) exit $?"; tail -8 gpurun_out/pytest_pg2.log
run() {
  name=$1; shift
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/bench_$name.json 2>gpurun_out/bench_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$name.json")); print("$name value", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "gemm ms/frame", round(d["kernel_breakdown"]["gemm_tc_kernel"]["ms_per_frame"],3))
except Exception as e: print("$name ERR", e); print(open("gpurun_out/bench_$name.err").read()[-2000:])
PY
}
run pg0
SMB_PGEMM=2 run pg2
SMB_PGEMM=3 run pg3
SMB_PGEMM=1 run pg1
for sp in 16 32 64; do
  SMB_DEC_SPLITS=$sp timeout 600 python bench.py --workload gated_decode --frames 128 --steps 1 --warmup 3 > gpurun_out/bench_dec_sp$sp.json 2>gpurun_out/bench_dec_sp$sp.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_dec_sp$sp.json")); print("decode splits $sp fps", round(d["value"],2), d["decode"]["tokens_per_s"], d["config"]["kv_len_end"])
except Exception as e: print("ERR", e); print(open("gpurun_out/bench_dec_sp$sp.err").read()[-1500:])
PY
done
