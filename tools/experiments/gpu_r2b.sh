#!/bin/bash
# round 2, second GPU pass: full GPU suite again; memory-only / compute-only timing builds; z-face cost experiment
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/r2b_pytest.log 2>&1
tail -4 gpurun_out/r2b_pytest.log
L=gpurun_out/r2b_diag.log; : > $L
for v in prod diag1 diag2; do
  lib=phonomena_b200/libphb200_$v.so; [ "$v" = prod ] && lib=phonomena_b200/libphb200.so
  for d in f64 f32; do
    echo "## $v $d" >> $L
    PHB200_LIB=$lib timeout 120 python tools/quick_bench.py --n 512 512 512 --dtype $d --kernel march --steps 20 2>&1 | tail -1 | cut -c1-200 >> $L
  done
done
echo "## prod f64 zfuse=1" >> $L
PHB_ZFUSE=1 timeout 120 python tools/quick_bench.py --n 512 512 512 --dtype f64 --kernel march --steps 20 2>&1 | tail -1 | cut -c1-200 >> $L
echo "## prod f64 zfuse nop (ZF code resident, never run; full z kernel)" >> $L
PHB_DEBUG_ZF_NOP=1 timeout 120 python tools/quick_bench.py --n 512 512 512 --dtype f64 --kernel march --steps 20 2>&1 | tail -1 | cut -c1-200 >> $L
echo "## prod f32 zfuse=0" >> $L
PHB_ZFUSE=0 timeout 120 python tools/quick_bench.py --n 512 512 512 --dtype f32 --kernel march --steps 20 2>&1 | tail -1 | cut -c1-200 >> $L
cat $L
