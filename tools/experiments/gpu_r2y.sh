#!/bin/bash
for d in f64 f32; do for z in 1 0; do for ch in 0 2 3 4 6 8 12 16; do
r=$(PHB_ZSPLIT=$z PHB_MARCH_CHUNKS=$ch timeout 60 python tools/quick_bench.py --n 256 256 256 --dtype $d --homog --steps 200 --warmup 20 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4), round(d['gcells'],1))")
echo "$d zsplit=$z chunks=$ch: $r"
done; done; done
