#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_variants.py tests/test_gpu_parity.py tests/test_gpu_edges.py tests/test_gpu_scale.py -m gpu -q -x ) > gpurun_out/r2o_pytest.log 2>&1; tail -2 gpurun_out/r2o_pytest.log
for z in 1 2 0; do
echo "## zsplit=$z f64"
PHB_ZSPLIT=$z timeout 120 python tools/quick_bench.py --n 512 512 512 --dtype f64 --kernel march --steps 30 --warmup 8 2>&1 | tail -1 | cut -c1-130
done
for ch in 1 3 4; do
echo "## chunks=$ch f64 (zsplit default)"
PHB_MARCH_CHUNKS=$ch timeout 120 python tools/quick_bench.py --n 512 512 512 --dtype f64 --kernel march --steps 30 --warmup 8 2>&1 | tail -1 | cut -c1-130
done
echo "## f32 zsplit=1 (3-way)"; PHB_ZSPLIT=1 timeout 120 python tools/quick_bench.py --n 512 512 512 --dtype f32 --kernel march --steps 30 --warmup 8 2>&1 | tail -1 | cut -c1-130
echo "## f32 default"; timeout 120 python tools/quick_bench.py --n 512 512 512 --dtype f32 --kernel march --steps 30 --warmup 8 2>&1 | tail -1 | cut -c1-130
