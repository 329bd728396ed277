#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/exp4.log
: > $L
export PHB_MARCH_RW=2
qb() { label=$1; d=$2; shift 2
  echo "## $label $d $*" >> $L
  env "$@" timeout 120 python tools/quick_bench.py --n 512 512 512 --dtype $d --kernel march --steps 20 2>&1 | tail -1 | cut -c1-200 >> $L
}
for v in u2 nso2; do
  for d in f64 f32; do qb $v $d PHB200_LIB=phonomena_b200/libphb200_$v.so; done
done
for d in f64 f32; do
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_step_march --launch-skip 3 --launch-count 1 \
  -f -o gpurun_out/r1c_march_$d python tools/quick_bench.py --n 512 512 512 --dtype $d --kernel march --steps 3 --warmup 2 > /dev/null 2>&1
done
cat $L
