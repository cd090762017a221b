#!/bin/bash
# measurement build of the library with the decode kernel's wait probes compiled in (not the product): SMB_LIB_PATH selects it
cd "$(dirname "$0")/../../streammind_b200/csrc" && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared \
  --expt-relaxed-constexpr -DSMB_DS_WAITPROBE ${EXTRA_DEFS} -I . api.cu -o ../libstreammind_b200_probe.so -lcudart
