#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_plugin.py -m gpu -q -x ) > gpurun_out/r2h_pytest.log 2>&1; tail -2 gpurun_out/r2h_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-disk --no-cpu --dtype f32 --e2e-fields ux,uy,uz > gpurun_out/r2h_bench_n1_f32.json 2> gpurun_out/r2h.err
python -c "import json;d=json.load(open('gpurun_out/r2h_bench_n1_f32.json'));e=d['e2e'];print('f32 value',d['value'],'e2e(3 comps)',e['value'],'run_ms',e['run_ms'],'loop',e['loop_ms'],'fin',e['writer_finish_ms'],'write',e['writer_write_ms'])"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --e2e-fields ux,uy,uz > gpurun_out/r2h_bench_n1_f64_3c.json 2>> gpurun_out/r2h.err
python -c "import json;d=json.load(open('gpurun_out/r2h_bench_n1_f64_3c.json'));e=d['e2e'];print('f64 value',d['value'],'e2e(3 comps)',e['value'],'disk',e.get('disk'))"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2h_launches.csv python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/r2h_bench_under_ncu.log 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open('gpurun_out/r2h_launches.csv')) if len(r)>5]
hdr=rows[0]; ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value'); 
agg=collections.OrderedDict()
for r in rows[1:]:
    n=r[ik].split('(')[0][:60]; agg.setdefault(n,[]).append(float(r[iv].replace(',','')))
for n,v in agg.items(): print('%-62s n=%3d mean=%.1f us'%(n,len(v),sum(v)/len(v)/1e3))
PY
