#!/bin/bash
# parity + 512^3 timing of the march kernel (fp64, fp32); used during kernel tuning
timeout 90 python -m pytest tests/test_gpu_parity.py -x -q -m gpu --timeout 30 2>&1 | tail -2
for d in f64 f32; do
  timeout 60 python tools/quick_bench.py --n 512 512 512 --dtype $d --kernel march --steps 20 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['dtype'], round(d['ms_per_step'],3), 'ms', round(d['gcells'],1), 'Gcell/s', round(d['frac_of_hbm_peak'],3))"
done
