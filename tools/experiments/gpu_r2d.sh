#!/bin/bash
# ncu --set full + source of the two stencil launches of one split step (face tile launch + the rest), fp64 512^3
mkdir -p gpurun_out
PHB_GRAPH=0 timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_step_march --launch-skip 8 --launch-count 2 \
  -f -o gpurun_out/r2d_split_f64 python tools/quick_bench.py --n 512 512 512 --dtype f64 --kernel march --steps 3 --warmup 5 > gpurun_out/r2d_ncu.log 2>&1
tail -3 gpurun_out/r2d_ncu.log
ls -la gpurun_out/*.ncu-rep
