#!/bin/bash
# Builds libstreammind_b200.so in-tree for sm_100a (cross-compiles without a GPU); ptxas statistics go to /tmp/smb_build.log.
set -e
cd "$(dirname "$0")"
OUT=../libstreammind_b200.so
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared \
     --expt-relaxed-constexpr -Xptxas -v -I. api.cu -o "$OUT" -lcudart 2> /tmp/smb_build.log || { cat /tmp/smb_build.log; exit 1; }
grep -E "error|warning" /tmp/smb_build.log | grep -v "ptxas info" | head -20 || true
echo "built $OUT"
