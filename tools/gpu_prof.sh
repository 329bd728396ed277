#!/bin/bash
# ncu --set full (+source) of the marching kernel for both dtypes -> gpurun_out/<tag>_march_{f64,f32}.ncu-rep
tag=${1:-r1c}
mkdir -p gpurun_out
for d in f64 f32; do
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_step_march --launch-skip 3 --launch-count 1 \
  -f -o gpurun_out/${tag}_march_$d python tools/quick_bench.py --n 512 512 512 --dtype $d --kernel march --steps 3 --warmup 2 > /dev/null 2>&1
done
ls -la gpurun_out/*.ncu-rep
