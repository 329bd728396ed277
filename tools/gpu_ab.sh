#!/bin/bash
# A/B timing of library variants at 512^3 (development aid):  tools/gpu_ab.sh [prod] [variant ...]
# variant = suffix of phonomena_b200/libphb200_<variant>.so built by tools/diag_build.sh; extra environment
# (PHB_MARCH_RW, PHB_ZFUSE, PHB_MARCH_CHUNKS, ...) is inherited.  Log: gpurun_out/ab.log
mkdir -p gpurun_out
L=gpurun_out/ab.log
: > $L
for v in "${@:-prod}"; do
  lib=phonomena_b200/libphb200_$v.so
  [ "$v" = prod ] && lib=phonomena_b200/libphb200.so
  PHB200_LIB=$lib timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edges.py tests/test_gpu_variants.py -x -q -m gpu 2>&1 | tail -1 >> $L
  for d in f64 f32; do
    echo "## $v $d" >> $L
    PHB200_LIB=$lib timeout 120 python tools/quick_bench.py --n 512 512 512 --dtype $d --kernel march --steps 20 2>&1 | tail -1 | cut -c1-160 >> $L
  done
done
cat $L
