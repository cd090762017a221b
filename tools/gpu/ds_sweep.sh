#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/ds_sweep.txt; : > $out
timeout 900 python -m pytest tests/test_llm_gpu.py tests/test_llm_fullsize_gpu.py -m gpu -q -x --timeout 600 -p no:cacheprovider > gpurun_out/pytest_llm.log 2>&1; echo "llm exit $?"
tail -5 gpurun_out/pytest_llm.log
echo "== no weight stream (SMB_DS_DBG=2)" >> $out
SMB_DS_DBG=2 timeout 300 python tools/decode_probe.py --layers 32 --ctx 2048 --phases 2>&1 | grep -E "^layers|CTA 0" >> $out
for args in "--ctx 2048 --phases" "--ctx 4096" "--ctx 8000" "--ctx 2048 --streams 2" "--ctx 2048 --streams 4"; do
  timeout 300 python tools/decode_probe.py --layers 32 $args 2>&1 | grep -E "^layers|CTA 0" >> $out
done
cat $out
