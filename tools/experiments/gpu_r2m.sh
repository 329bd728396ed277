#!/bin/bash
mkdir -p gpurun_out
for sp in 0 1 2 3; do
echo "## self_push=$sp"
PHB_DEBUG_SELF_PUSH=$sp timeout 120 python tools/quick_bench.py --n 512 512 512 --dtype f64 --kernel march --steps 30 --warmup 8 2>&1 | tail -1 | cut -c1-130
done
