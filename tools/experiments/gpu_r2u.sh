#!/bin/bash
( timeout 600 python -m pytest tests/test_gpu_plugin.py -m gpu -q -x ) 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
for ch in 1 3 4 6; do
echo "## f32 chunks=$ch"
PHB_MARCH_CHUNKS=$ch timeout 120 python tools/quick_bench.py --n 512 512 512 --dtype f32 --kernel march --steps 30 --warmup 8 2>&1 | tail -1 | cut -c60-130
done
