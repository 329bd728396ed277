#!/bin/bash
# round 2, first GPU pass: full GPU test suite, both bench arms, z-face A/B, launch list
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2a_pytest.log 2>&1
tail -5 gpurun_out/r2a_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench_n1.json 2> gpurun_out/r2a_bench_n1.err
tail -c 1500 gpurun_out/r2a_bench_n1.json; tail -3 gpurun_out/r2a_bench_n1.err
for z in 0 1; do
  PHB_ZFUSE=$z timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu > gpurun_out/r2a_bench_zf$z.json 2>> gpurun_out/r2a_bench_n1.err
  python -c "import json;d=json.load(open('gpurun_out/r2a_bench_zf$z.json'));print('zfuse=$z ms/step',d['ms_per_step'],'kernel',d['roofline']['kernel_ms_per_step'])"
done
PHB_GRAPH=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu > gpurun_out/r2a_bench_nograph.json 2>> gpurun_out/r2a_bench_n1.err
python -c "import json;d=json.load(open('gpurun_out/r2a_bench_nograph.json'));print('nograph ms/step',d['ms_per_step'])"
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2a_bench_ref.json 2> gpurun_out/r2a_bench_ref.err
tail -c 2500 gpurun_out/r2a_bench_ref.json; tail -3 gpurun_out/r2a_bench_ref.err
