#!/bin/bash
# N GPUs: slab parity check, weak + strong bench lines
N=${1:-2}
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
if [ "$N" = 1 ]; then
  ( time timeout 1200 python -m pytest tests/test_gpu_periodic.py tests/test_gpu_long.py -m gpu -q -s ) > gpurun_out/r2f_pytest.log 2>&1
  tail -8 gpurun_out/r2f_pytest.log
  timeout 900 python bench.py --scaling strong --steps 20 --warmup 5 --no-cpu --no-disk > gpurun_out/r2_strong_n1.json 2> gpurun_out/r2_strong_n1.err
  python -c "import json;d=json.load(open('gpurun_out/r2_strong_n1.json'));print('strong n1',d['value'],d['ms_per_step'],d['e2e']['value'])"
  exit 0
fi
timeout 900 $T tools/multi_check.py > gpurun_out/r2_multicheck_n$N.log 2>&1; echo "multi_check rc=$?"; grep -c '"ranks_identical": '$N gpurun_out/r2_multicheck_n$N.log; grep selfcheck gpurun_out/r2_multicheck_n$N.log | cut -c1-200; tail -2 gpurun_out/r2_multicheck_n$N.log | cut -c1-300
timeout 900 $T bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_weak_n$N.json 2> gpurun_out/r2_weak_n$N.err; echo "weak rc=$?"
timeout 900 $T bench.py --gpus $N --scaling strong --steps 20 --warmup 5 > gpurun_out/r2_strong_n$N.json 2> gpurun_out/r2_strong_n$N.err; echo "strong rc=$?"
for f in weak strong; do python -c "
import json;d=json.loads(open('gpurun_out/r2_${f}_n$N.json').read().strip().splitlines()[-1]);print('$f n$N value',d['value'],'ms',d['ms_per_step'],'kernel',d['roofline']['kernel_ms_per_step'],'e2e',d['e2e']['value'],'parity',d.get('parity'))"; done
tail -3 gpurun_out/r2_strong_n$N.err
