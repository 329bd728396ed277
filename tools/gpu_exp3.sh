#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/exp3.log
: > $L
qb() { label=$1; d=$2; shift 2
  echo "## $label $d $*" >> $L
  env "$@" timeout 120 python tools/quick_bench.py --n 512 512 512 --dtype $d --kernel march --steps 20 2>&1 | tail -1 | cut -c1-200 >> $L
}
for rw in 2 1; do
  echo "#### RW=$rw" >> $L
  PHB_MARCH_RW=$rw timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py -x -q -m gpu 2>&1 | tail -3 >> $L
  for d in f64 f32; do qb rw$rw $d PHB_MARCH_RW=$rw; done
done
cat $L
