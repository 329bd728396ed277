#!/bin/bash
# Round artefacts: ncu full captures (both dtypes), ncu launch list of the bench command, bench lines (N=1 + reference arm)
tag=${1:-r1e}
mkdir -p gpurun_out
bash tools/gpu_prof.sh $tag > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
   python bench.py --steps 5 --warmup 3 --no-e2e > gpurun_out/${tag}_bench_under_ncu.log 2>&1
python bench.py --extra > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err
python bench.py --dtype f32 > gpurun_out/${tag}_bench_n1_f32.json 2>> gpurun_out/${tag}_bench_n1.err
python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/${tag}_bench_reference_arm.json 2>> gpurun_out/${tag}_bench_n1.err
ls -la gpurun_out | tail -12
tail -c 600 gpurun_out/${tag}_bench_n1.json
