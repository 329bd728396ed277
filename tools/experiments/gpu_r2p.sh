#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_variants.py tests/test_gpu_parity.py tests/test_gpu_edges.py tests/test_gpu_scale.py tests/test_gpu_periodic.py -m gpu -q -x ) > gpurun_out/r2p_pytest.log 2>&1; tail -2 gpurun_out/r2p_pytest.log
for z in 1 2 0; do
echo "## zsplit=$z f64"
PHB_ZSPLIT=$z timeout 120 python tools/quick_bench.py --n 512 512 512 --dtype f64 --kernel march --steps 30 --warmup 8 2>&1 | tail -1 | cut -c1-130
done
echo "## f32 zsplit=1 (3-way)"; PHB_ZSPLIT=1 timeout 120 python tools/quick_bench.py --n 512 512 512 --dtype f32 --kernel march --steps 30 --warmup 8 2>&1 | tail -1 | cut -c1-130
echo "## f32 zsplit=2"; PHB_ZSPLIT=2 timeout 120 python tools/quick_bench.py --n 512 512 512 --dtype f32 --kernel march --steps 30 --warmup 8 2>&1 | tail -1 | cut -c1-130
echo "## f32 default"; timeout 120 python tools/quick_bench.py --n 512 512 512 --dtype f32 --kernel march --steps 30 --warmup 8 2>&1 | tail -1 | cut -c1-130
timeout 600 python bench.py --steps 20 --warmup 5 --no-disk --no-cpu > gpurun_out/r2p_bench_n1.json 2> gpurun_out/r2p.err
python -c "import json;d=json.load(open('gpurun_out/r2p_bench_n1.json'));print('bench value',d['value'],'ms',d['ms_per_step'],'kernel',d['roofline']['kernel_ms_per_step'],'frac',d['roofline']['frac'],'step_frac',d['roofline']['step_frac'],'e2e',d['e2e']['value'],'launches',d['gpu_launches'])"
