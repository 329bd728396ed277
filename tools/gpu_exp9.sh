#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/exp9.log
: > $L
qb() { label=$1; d=$2; shift 2
  echo "## $label $d $*" >> $L
  env "$@" timeout 120 python tools/quick_bench.py --n 512 512 512 --dtype $d --kernel march --steps 20 2>&1 | tail -1 | cut -c1-150 >> $L
}
qb k0 f64 A=1
qb k0 f32 A=1
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py tests/test_gpu_variants.py tests/test_gpu_plugin.py -x -q -m gpu 2>&1 | tail -2 >> $L
python bench.py --steps 50 > gpurun_out/bench_tmp.json 2> gpurun_out/bench_tmp.err
python -c "
import json
d=json.load(open('gpurun_out/bench_tmp.json'))
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e'])" >> $L
cat $L
