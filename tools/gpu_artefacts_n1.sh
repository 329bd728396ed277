#!/bin/bash
# final single-GPU artefacts of round 2: long run, bench lines (f64 default, f32, reference arm), launch list
# (the ncu --set full captures are separate calls, tools/gpu_ncu_full.sh f64|f32: gpurun copies back at most 64 MiB per call)
mkdir -p gpurun_out
timeout 1500 python bench.py --steps 10000 --warmup 5 --no-cpu --no-disk > gpurun_out/r2_long_n1.json 2> gpurun_out/r2_long_n1.err
python -c "import json;d=json.load(open('gpurun_out/r2_long_n1.json'));print('long n1',d['value'],d['ms_per_step'],d['clocks'],d['e2e']['value'])"
timeout 900 python bench.py --steps 20 --warmup 5 --extra > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
python -c "import json;d=json.load(open('gpurun_out/r2_bench_n1.json'));print('n1',d['value'],d['ms_per_step'],d['e2e']['value'],d['e2e'].get('disk',{}).get('value'),d['cpu_baseline']['value'])"
timeout 900 python bench.py --steps 20 --warmup 5 --dtype f32 --no-cpu --no-disk > gpurun_out/r2_bench_n1_f32.json 2>> gpurun_out/r2_bench_n1.err
timeout 900 python bench.py --steps 20 --warmup 5 --dtype f32 --no-cpu --no-disk --e2e-fields ux,uy,uz > gpurun_out/r2_bench_n1_f32_3c.json 2>> gpurun_out/r2_bench_n1.err
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu --no-disk --e2e-fields ux,uy,uz > gpurun_out/r2_bench_n1_f64_3c.json 2>> gpurun_out/r2_bench_n1.err
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench_reference_arm.json 2> gpurun_out/r2_bench_ref.err
timeout 900 python bench.py --scaling strong --steps 20 --warmup 5 --no-cpu --no-disk > gpurun_out/r2_strong_n1.json 2> gpurun_out/r2_strong_n1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2_bench_under_ncu.log 2>&1
ls -la gpurun_out/r2_launches.csv
