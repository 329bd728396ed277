#!/bin/bash
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q -x ) > gpurun_out/r2g_pytest.log 2>&1
tail -5 gpurun_out/r2g_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-disk --no-cpu > gpurun_out/r2g_bench_n1.json 2> gpurun_out/r2g_bench_n1.err
python -c "import json;d=json.load(open('gpurun_out/r2g_bench_n1.json'));print('bench value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'])"
timeout 600 python bench.py --steps 20 --warmup 5 --no-disk --no-cpu --dtype f32 --e2e-fields ux,uy,uz > gpurun_out/r2g_bench_n1_f32.json 2>> gpurun_out/r2g_bench_n1.err
python -c "import json;d=json.load(open('gpurun_out/r2g_bench_n1_f32.json'));print('f32 value',d['value'],'ms',d['ms_per_step'],'e2e(3 comps)',d['e2e']['value'])"
