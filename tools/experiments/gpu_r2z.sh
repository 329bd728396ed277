#!/bin/bash
for n in "128 512 512" "256 512 512" "384 384 384" "320 320 320"; do for d in f64 f32; do for z in 1 0; do
r=$(PHB_ZSPLIT=$z timeout 60 python tools/quick_bench.py --n $n --dtype $d --steps 100 --warmup 20 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4), round(d['gcells'],1))")
echo "$n $d zsplit=$z: $r"
done; done; done
