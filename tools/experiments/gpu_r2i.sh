#!/bin/bash
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q -x ) > gpurun_out/r2i_pytest.log 2>&1; tail -5 gpurun_out/r2i_pytest.log
for ff in 1 0; do
PHB_FACES_FUSED=$ff timeout 600 python bench.py --steps 20 --warmup 5 --no-disk --no-cpu --no-e2e > gpurun_out/r2i_bench_ff$ff.json 2>> gpurun_out/r2i.err
python -c "import json;d=json.load(open('gpurun_out/r2i_bench_ff$ff.json'));print('faces_fused=$ff value',d['value'],'ms',d['ms_per_step'],'launches',d['gpu_launches'])"
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-disk --no-cpu --dtype f32 --e2e-fields ux,uy,uz > gpurun_out/r2i_bench_n1_f32.json 2>> gpurun_out/r2i.err
python -c "import json;d=json.load(open('gpurun_out/r2i_bench_n1_f32.json'));e=d['e2e'];print('f32 value',d['value'],'e2e(3 comps)',e['value'],'run_ms',e['run_ms'],'loop',e['loop_ms'],'fin',e['writer_finish_ms'],'write',e['writer_write_ms'])"
