#!/bin/bash
mkdir -p gpurun_out
echo "## config 2: homogeneous 256^3"
for d in f64 f32; do timeout 120 python tools/quick_bench.py --n 256 256 256 --dtype $d --homog --steps 200 --warmup 20 2>&1 | tail -1 | cut -c1-200; done
echo "## config 1: default.json-sized grid through the plugin"
timeout 300 python tools/config1.py 2>&1 | tail -6
echo "## fp32 long run"
timeout 900 python bench.py --steps 10000 --warmup 5 --dtype f32 --no-cpu --no-disk > gpurun_out/r2_long_n1_f32.json 2>/dev/null
python -c "import json;d=json.load(open('gpurun_out/r2_long_n1_f32.json'));print('long f32',d['value'],d['ms_per_step'],d['clocks'],d['e2e']['value'])"
