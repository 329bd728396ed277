#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/exp8.log
: > $L
qb() { label=$1; d=$2; shift 2
  echo "## $label $d $*" >> $L
  env "$@" timeout 120 python tools/quick_bench.py --n 512 512 512 --dtype $d --kernel march --steps 20 2>&1 | tail -1 | cut -c1-150 >> $L
}
for ch in 2 3 4 6; do for z in 0 1; do qb x f64 PHB_ZFUSE=$z PHB_MARCH_CHUNKS=$ch; done; done
for ch in 2 4 6 8; do qb x f32 PHB_ZFUSE=1 PHB_MARCH_CHUNKS=$ch; done
cat $L
