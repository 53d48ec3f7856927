#!/bin/bash
# Development helper: builds cfg4-only variants of the library (CKKS_ONLY_CFG4: 64-bit words, A = 8, lazy8 -- about
# 35 s each instead of 3.5 min) with different -D switches into variants/, to be compared on the GPU with
#   CKKS_B200_LIB=variants/libckks_<name>.so python bench.py --no-ntt --no-chain --no-single-thread ...
# usage: tools/variants.sh name1 "-DFOO=1 -DBAR=2" name2 "..." ...
set -e
cd "$(dirname "$0")/.."
mkdir -p variants
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  ( /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared \
      -DCKKS_ONLY_CFG4 $flags -o variants/libckks_$name.so toy-heaan-ckks_b200/csrc/ckks_b200.cu -lcudart > variants/build_$name.log 2>&1 \
      && echo "built $name" || echo "FAILED $name" ) &
done
wait
