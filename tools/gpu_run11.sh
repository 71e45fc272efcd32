#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_strips_sigma_relax_gpu.py -q -m gpu > gpurun_out/r2_11_strips.log 2>&1
echo "strips rc=$?" >> gpurun_out/r2_11_strips.log
timeout 1800 python -m pytest tests -q -m gpu -x --deselect tests/test_parity_at_baseline_sizes_gpu.py --deselect tests/test_strips_sigma_relax_gpu.py > gpurun_out/r2_11_gputests.log 2>&1
echo "gputests rc=$?" >> gpurun_out/r2_11_gputests.log
grep -E "^E  .*Error|^E  .*assert|passed|failed|^FAILED|rc=" gpurun_out/r2_11_strips.log gpurun_out/r2_11_gputests.log | cut -c1-500 | head -40
