#!/bin/bash
# N GPUs: slab parity (multi_check), weak + strong bench lines; LONG=1 adds the 10 000-step record (config #5)
N=${1:-2}; LONG=${2:-0}
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
show() { python -c "
import json,sys
d=json.loads(open('$1').read().strip().splitlines()[-1]); e=d.get('e2e') or {}
print('$2 value %.2f ms %.4f kernel %.4f e2e %.2f parity %s clocks %s' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_step'], e.get('value',0), (d.get('parity') or {}).get('slabs_bit_identical'), d['clocks']))"; }
if [ "$N" = 1 ]; then
  timeout 1500 python bench.py --steps 10000 --warmup 5 --no-cpu --no-disk > gpurun_out/r2_long_n1.json 2> gpurun_out/r2_long_n1.err; show gpurun_out/r2_long_n1.json "long n1"
  exit 0
fi
timeout 900 $T tools/multi_check.py > gpurun_out/r2_multicheck_n$N.log 2>&1; echo "multi_check rc=$?"; grep -c '"ranks_identical": '$N gpurun_out/r2_multicheck_n$N.log; grep -E "selfcheck|plugin over" gpurun_out/r2_multicheck_n$N.log | cut -c1-220
timeout 900 $T bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_weak_n$N.json 2> gpurun_out/r2_weak_n$N.err; echo "weak rc=$?"; show gpurun_out/r2_weak_n$N.json "weak n$N"
timeout 900 $T bench.py --gpus $N --scaling strong --steps 20 --warmup 5 > gpurun_out/r2_strong_n$N.json 2> gpurun_out/r2_strong_n$N.err; echo "strong rc=$?"; show gpurun_out/r2_strong_n$N.json "strong n$N"
if [ "$LONG" = 1 ]; then
  timeout 1500 $T bench.py --gpus $N --steps 10000 --warmup 5 > gpurun_out/r2_long_n$N.json 2> gpurun_out/r2_long_n$N.err; echo "long rc=$?"; show gpurun_out/r2_long_n$N.json "long n$N"
fi
grep -v "OMP_NUM_THREADS\|^\*\*\*" gpurun_out/r2_strong_n$N.err | tail -3
