#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_reference_shaders_parity_gpu.py -q -m gpu -k "occ or _do" > gpurun_out/r2_6_occ.log 2>&1
echo "occ rc=$?" >> gpurun_out/r2_6_occ.log
grep -E "^E  |passed|failed|Error" gpurun_out/r2_6_occ.log | cut -c1-600 | head -40
