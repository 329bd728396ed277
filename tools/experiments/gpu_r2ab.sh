#!/bin/bash
for d in f64 f32; do
for env in "A=1" "PHB_MARCH_R=8" "PHB_MARCH_R=8 PHB_MARCH_CHUNKS=4" "PHB_MARCH_R=8 PHB_MARCH_CHUNKS=6" "PHB_MARCH_R=8 PHB_MARCH_CHUNKS=8" "PHB_MARCH_RW=1" "PHB_MARCH_RW=1 PHB_MARCH_CHUNKS=6"; do
r=$(env $env timeout 60 python tools/quick_bench.py --n 256 256 256 --dtype $d --homog --steps 200 --warmup 20 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],4), round(d['gcells'],1))")
echo "256^3 $d [$env]: $r"
done; done
