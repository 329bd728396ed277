#!/bin/bash
# ncu --set full of the stencil launches of one step: fp64 = three launches (split step), fp32 = one
mkdir -p gpurun_out
PHB_GRAPH=0 timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_step_march --launch-skip 9 --launch-count 3 \
  -f -o gpurun_out/r2_march_f64 python tools/quick_bench.py --n 512 512 512 --dtype f64 --kernel march --steps 3 --warmup 5 > gpurun_out/r2_ncu_f64.log 2>&1
PHB_GRAPH=0 timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_step_march --launch-skip 6 --launch-count 1 \
  -f -o gpurun_out/r2_march_f32 python tools/quick_bench.py --n 512 512 512 --dtype f32 --kernel march --steps 3 --warmup 5 > gpurun_out/r2_ncu_f32.log 2>&1
ls -la gpurun_out/r2_march_*.ncu-rep
