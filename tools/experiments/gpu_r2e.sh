#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_periodic.py tests/test_gpu_plugin.py tests/test_gpu_parity.py -m gpu -q ) > gpurun_out/r2e_pytest.log 2>&1
tail -6 gpurun_out/r2e_pytest.log
bash tools/gpu_r2d.sh
