#!/bin/bash
# builds sclane_b200/lib/variant_<name>.so with extra nvcc flags:  tools/build_variant.sh <name> -DSCLANE_X=1 ...
cd "$(dirname "$0")/.."
name=$1; shift
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --shared -Xcompiler -fPIC -Xcompiler -pthread \
  -I include "$@" sclane_b200/csrc/sclane_api.cu sclane_b200/csrc/sclane_narrow.cpp -o sclane_b200/lib/variant_$name.so && echo built variant_$name
