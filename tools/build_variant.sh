#!/bin/bash
# build an experiment variant of libwfe.so into exp_so/: tools/build_variant.sh <name> [-DFLAG ...]
set -e
name=$1; shift
mkdir -p "$(dirname "$0")/../exp_so"
cd "$(dirname "$0")/../asr-finetune_b200/csrc"
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-O3 -DWFE_EXP_MINIMAL "$@" \
  -shared -o ../../exp_so/libwfe_$name.so wfe_api.cu -lcudart
echo built exp_so/libwfe_$name.so
