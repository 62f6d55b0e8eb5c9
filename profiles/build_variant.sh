#!/bin/bash
# A/B builds of the library: profiles/build_variant.sh NAME "-DFLAG=.. -DFLAG2=.."  ->  pylians_b200/lib/variants/NAME.so
# (deposit_tiled.cu recompiled with the flags, the other objects reused).  The run scripts copy a variant over
# pylians_b200/lib/libpylians_b200.so on the GPU box's scratch copy of the repo.
set -e
cd "$(dirname "$0")/.."
mkdir -p pylians_b200/lib/variants
nvcc $2 -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-O3 --expt-relaxed-constexpr -Xcudafe --diag_suppress=177 \
  -c pylians_b200/csrc/deposit_tiled.cu -o /tmp/dt_$1.o
objs=$(ls pylians_b200/lib/obj/*.o | grep -v deposit_tiled.o)
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o pylians_b200/lib/variants/$1.so /tmp/dt_$1.o $objs -lcufft -Xlinker -rpath,/usr/local/cuda/lib64
echo built pylians_b200/lib/variants/$1.so
