#!/bin/bash
# usage: gpu_exp5.sh <variant>...   (prod = the product library); parity with RW=2 and timing for RW=2 / RW=1
mkdir -p gpurun_out
L=gpurun_out/exp5.log
: > $L
qb() { label=$1; d=$2; shift 2
  echo "## $label $d $*" >> $L
  env "$@" timeout 120 python tools/quick_bench.py --n 512 512 512 --dtype $d --kernel march --steps 20 2>&1 | tail -1 | cut -c1-150 >> $L
}
for v in "$@"; do
  lib=phonomena_b200/libphb200_$v.so
  [ "$v" = prod ] && lib=phonomena_b200/libphb200.so
  echo "#### $v" >> $L
  for rw in 2 1; do
    PHB_MARCH_RW=$rw PHB200_LIB=$lib timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py -x -q -m gpu 2>&1 | tail -1 >> $L
    for d in f64 f32; do qb $v $d PHB200_LIB=$lib PHB_MARCH_RW=$rw; done
  done
done
cat $L
