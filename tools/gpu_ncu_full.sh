#!/bin/bash
# ncu --set full of the three stencil launches of one step (split step), one precision per call (64 MiB copy-back limit)
d=${1:-f64}
mkdir -p gpurun_out
PHB_GRAPH=0 timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_step_march --launch-skip 9 --launch-count 3 \
  -f -o gpurun_out/r2_march_$d python tools/quick_bench.py --n 512 512 512 --dtype $d --kernel march --steps 3 --warmup 5 > gpurun_out/r2_ncu_$d.log 2>&1
ls -la gpurun_out/r2_march_$d.ncu-rep
