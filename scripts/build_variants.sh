#!/bin/bash
# Build tuning variants of the library at the repo root: scripts/build_variants.sh name="-DFLAG=.. -DFLAG2=.." ...
# (variants_<name>.so; bench them on the GPU box with scripts/gpu_variants.sh)
cd "$(dirname "$0")/../decompress_b200/csrc"
for spec in "$@"; do
  name="${spec%%=*}"; flags="${spec#*=}"
  /usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC --cudart static \
    $flags -shared -o ../../variants_$name.so *.cu 2>&1 | grep -v "^$" &
done
wait
ls -la ../../variants_*.so
