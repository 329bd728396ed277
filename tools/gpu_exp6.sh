#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/exp6.log
: > $L
qb() { label=$1; d=$2; shift 2
  echo "## $label $d $*" >> $L
  env "$@" timeout 120 python tools/quick_bench.py --n 512 512 512 --dtype $d --kernel march --steps 20 2>&1 | tail -1 | cut -c1-150 >> $L
}
for v in "$@"; do
  lib=phonomena_b200/libphb200_$v.so
  [ "$v" = prod ] && lib=phonomena_b200/libphb200.so
  for d in f64 f32; do qb $v $d PHB200_LIB=$lib; done
done
python bench.py --steps 50 > gpurun_out/bench_tmp.json 2> gpurun_out/bench_tmp.err
python -c "
import json
d=json.load(open('gpurun_out/bench_tmp.json'))
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e'])" >> $L
cat $L
