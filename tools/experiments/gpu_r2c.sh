#!/bin/bash
# round 2, third GPU pass: split-step launch (face tile on its own stream) -- parity of the variants, A/B timing
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_variants.py tests/test_gpu_parity.py tests/test_gpu_edges.py tests/test_gpu_plugin.py -m gpu -q ) > gpurun_out/r2c_pytest.log 2>&1
tail -4 gpurun_out/r2c_pytest.log
L=gpurun_out/r2c_ab.log; : > $L
run() {  # label, lib, dtype, env...
  local label=$1 lib=$2 d=$3; shift 3
  echo "## $label $d" >> $L
  env "$@" PHB200_LIB=$lib timeout 120 python tools/quick_bench.py --n 512 512 512 --dtype $d --kernel march --steps 30 --warmup 8 2>&1 | tail -1 | cut -c1-170 >> $L
}
P=phonomena_b200/libphb200.so; U=phonomena_b200/libphb200_unroll1.so
for d in f64 f32; do
  run "split (default)" $P $d A=1
  run "zsplit=0" $P $d PHB_ZSPLIT=0
  run "zfuse=0" $P $d PHB_ZFUSE=0
  run "unroll1 split" $U $d A=1
  run "unroll1 zfuse=0" $U $d PHB_ZFUSE=0
done
cat $L
timeout 600 python bench.py --steps 20 --warmup 5 --no-disk > gpurun_out/r2c_bench_n1.json 2> gpurun_out/r2c_bench_n1.err
python -c "import json;d=json.load(open('gpurun_out/r2c_bench_n1.json'));print('bench value',d['value'],'ms',d['ms_per_step'],'kernel',d['roofline']['kernel_ms_per_step'],'e2e',d['e2e']['value'],'launches',d['gpu_launches'])"; tail -3 gpurun_out/r2c_bench_n1.err
