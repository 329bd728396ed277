#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/exp7.log
: > $L
qb() { label=$1; d=$2; shift 2
  echo "## $label $d $*" >> $L
  env "$@" timeout 120 python tools/quick_bench.py --n 512 512 512 --dtype $d --kernel march --steps 20 2>&1 | tail -1 | cut -c1-150 >> $L
}
for z in 0 1; do for d in f64 f32; do qb zfuse$z $d PHB_ZFUSE=$z; done; done
PHB_ZFUSE=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py tests/test_gpu_plugin.py -x -q -m gpu 2>&1 | tail -2 >> $L
cat $L
