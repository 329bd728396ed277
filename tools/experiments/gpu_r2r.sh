#!/bin/bash
for pr in 256 128 64 0; do
for d in f64 f32; do
echo "## promo=$pr $d"
PHB_TMA_PROMO=$pr timeout 120 python tools/quick_bench.py --n 512 512 512 --dtype $d --kernel march --steps 30 --warmup 8 2>&1 | tail -1 | cut -c1-130
done; done
