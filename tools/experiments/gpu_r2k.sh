#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 --no-disk --no-cpu --dtype f32 --e2e-fields ux,uy,uz > gpurun_out/r2k_bench_n1_f32.json 2>> gpurun_out/r2k.err
python -c "import json;d=json.load(open('gpurun_out/r2k_bench_n1_f32.json'));e=d['e2e'];print('f32 value',d['value'],'e2e(3 comps)',e['value'],'run_ms',e['run_ms'],'loop',e['loop_ms'],'fin',e['writer_finish_ms'],'write',e['writer_write_ms'],'init',e['init_s'])"
for ff in 1 0; do
PHB_GRAPH=0 PHB_FACES_FUSED=$ff timeout 600 python bench.py --steps 20 --warmup 5 --no-disk --no-cpu --no-e2e > gpurun_out/r2k_bench_ff$ff.json 2>> gpurun_out/r2k.err
python -c "import json;d=json.load(open('gpurun_out/r2k_bench_ff$ff.json'));print('nograph faces_fused=$ff value',d['value'],'ms',d['ms_per_step'],'launches',d['gpu_launches'])"
done
