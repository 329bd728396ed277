#!/bin/bash
# one-off experiment batch (development): where does K1's time go?
mkdir -p gpurun_out
L=gpurun_out/exp1.log
: > $L
qb() { # label dtype [ENV=VAL ...]
  label=$1; d=$2; shift 2
  echo "## $label $d $*" >> $L
  env "$@" timeout 120 python tools/quick_bench.py --n 512 512 512 --dtype $d --kernel march --steps 20 2>&1 | tail -1 >> $L
}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $L
for d in f64 f32; do
  qb base $d A=1
  qb nst3 $d PHB_MARCH_NST=3
  qb chunks1 $d PHB_MARCH_CHUNKS=1
  qb chunks4 $d PHB_MARCH_CHUNKS=4
  qb memonly $d PHB200_LIB=phonomena_b200/libphb200_mem.so
  qb cmponly $d PHB200_LIB=phonomena_b200/libphb200_cmp.so
done
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_step_march --launch-skip 3 --launch-count 1 \
  -f -o gpurun_out/r1b_march_f64 python tools/quick_bench.py --n 512 512 512 --dtype f64 --kernel march --steps 3 --warmup 2 >> $L 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_step_march --launch-skip 3 --launch-count 1 \
  -f -o gpurun_out/r1b_march_f32 python tools/quick_bench.py --n 512 512 512 --dtype f32 --kernel march --steps 3 --warmup 2 >> $L 2>&1
ls -la gpurun_out >> $L
cat $L
