#!/bin/bash
for z in 0 1 2; do for pr in 256 64 0; do
echo "## f32 zsplit=$z promo=$pr"
PHB_ZSPLIT=$z PHB_TMA_PROMO=$pr timeout 120 python tools/quick_bench.py --n 512 512 512 --dtype f32 --kernel march --steps 30 --warmup 8 2>&1 | tail -1 | cut -c60-130
done; done
for ch in 1 3 4; do
echo "## f64 chunks=$ch"
PHB_MARCH_CHUNKS=$ch timeout 120 python tools/quick_bench.py --n 512 512 512 --dtype f64 --kernel march --steps 30 --warmup 8 2>&1 | tail -1 | cut -c60-130
done
