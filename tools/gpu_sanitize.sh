#!/bin/bash
# compute-sanitizer memcheck over the small-grid parity tests that exercise every kernel of the step (marching kernel in all
# its split parts, faces, periodic / Bloch fix-up, recorder, probes); summary -> gpurun_out/r2_memcheck.log
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 77 --print-limit 20 \
  python -m pytest tests/test_gpu_periodic.py tests/test_gpu_edges.py tests/test_gpu_plugin.py -m gpu -q -x -k "not 10k and not cancel" > gpurun_out/r2_memcheck.log 2>&1
echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|passed|failed|Invalid|Error" gpurun_out/r2_memcheck.log | tail -8
