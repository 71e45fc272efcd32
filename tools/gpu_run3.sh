#!/bin/bash
mkdir -p gpurun_out
python tools/mirror_rates.py 960 540 3 > gpurun_out/r2_3_mirror_rates.log 2>&1
timeout 900 python -m pytest tests/test_parity_at_baseline_sizes_gpu.py -q -m gpu > gpurun_out/r2_3_bigparity.log 2>&1
echo "bigparity rc=$?" >> gpurun_out/r2_3_bigparity.log
timeout 1500 python -m pytest tests -x -q -m gpu --deselect tests/test_parity_at_baseline_sizes_gpu.py > gpurun_out/r2_3_gputests.log 2>&1
echo "gputests rc=$?" >> gpurun_out/r2_3_gputests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_3_bench_reblur.json 2> gpurun_out/r2_3_bench_reblur.err
cat gpurun_out/r2_3_mirror_rates.log; tail -n 4 gpurun_out/r2_3_bigparity.log gpurun_out/r2_3_gputests.log
