#!/bin/bash
( timeout 900 python -m pytest tests -m gpu -q -x ) 2>&1 | tail -2
for n in "512 512 512" "256 256 256" "128 512 512" "320 320 320"; do for d in f64 f32; do
r=$(timeout 60 python tools/quick_bench.py --n $n --dtype $d --steps 60 --warmup 20 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4), round(d['gcells'],1), d['launches'])")
echo "$n $d auto: $r"
done; done
