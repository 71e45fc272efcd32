#!/bin/bash
# Round-2 GPU batch 2: geometry plane + unified mirror predicate + packed history filter
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_at_baseline_sizes_gpu.py -q -m gpu > gpurun_out/r2_2_bigparity.log 2>&1
echo "bigparity rc=$?" >> gpurun_out/r2_2_bigparity.log
timeout 1500 python -m pytest tests -x -q -m gpu --deselect tests/test_parity_at_baseline_sizes_gpu.py > gpurun_out/r2_2_gputests.log 2>&1
echo "gputests rc=$?" >> gpurun_out/r2_2_gputests.log
bash tools/run_variants.sh > gpurun_out/r2_2_variants.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_2_bench_reblur.json 2> gpurun_out/r2_2_bench_reblur.err
timeout 300 python bench.py --denoiser relax --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_2_bench_relax.json 2> gpurun_out/r2_2_bench_relax.err
tail -n 3 gpurun_out/r2_2_bigparity.log gpurun_out/r2_2_gputests.log; cat gpurun_out/r2_2_variants.log
