#!/bin/bash
# experiment builds of the probe: exp_build.sh <tag> <nvcc -D flags...>  ->  ../libgavisunk_b200_<tag>.so
# (K = 20 only; every other object comes from the regular build)
set -e
cd "$(dirname "$0")"
tag=$1; shift
mkdir -p build_exp
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --extended-lambda -Xcompiler -fPIC \
  -Xcudafe --diag_suppress=177 -DGVS_PROBE_ONLY_K=20 "$@" -Xptxas -v -c probe.cu -o build_exp/probe_$tag.o 2> build_exp/probe_$tag.log
objs=$(ls build/*.o | grep -v build/probe.o)
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libgavisunk_b200_$tag.so $objs build_exp/probe_$tag.o -lz
grep -A1 "k_probe2ILi20ELb0ELb0" build_exp/probe_$tag.log | grep -E "Used|spill" | head -4
