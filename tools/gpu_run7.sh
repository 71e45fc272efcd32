#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_reference_shaders_parity_gpu.py -q -m gpu -k "dynamic" > gpurun_out/r2_7_dynres.log 2>&1
echo "dynres rc=$?" >> gpurun_out/r2_7_dynres.log
grep -E "^E  .*Error|passed|failed|^FAILED" gpurun_out/r2_7_dynres.log | cut -c1-500 | head -40
