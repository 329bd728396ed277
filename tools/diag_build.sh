#!/bin/bash
# Development aid: build timing variants of the library next to the product one (never shipped / never loaded by default).
#   tools/diag_build.sh <name> <extra nvcc flags...>   ->  phonomena_b200/libphb200_<name>.so
# Load with PHB200_LIB=phonomena_b200/libphb200_<name>.so (honoured by phonomena_b200/_lib.py).
set -e
cd "$(dirname "$0")/../phonomena_b200/csrc"
name=$1; shift
nvcc -O3 -std=c++20 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-Wall,-Wno-unused-function \
     --expt-relaxed-constexpr "$@" -shared -o ../libphb200_${name}.so phb200.cu -lcudart -ldl
