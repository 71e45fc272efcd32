#!/bin/bash
# Round-2 GPU batch 1: parity at the BASELINE sizes, bench lines, full GPU suite
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_1_gpus.txt 2>&1
nproc >> gpurun_out/r2_1_gpus.txt; cat /sys/fs/cgroup/cpu.max >> gpurun_out/r2_1_gpus.txt 2>&1
timeout 900 python -m pytest tests/test_parity_at_baseline_sizes_gpu.py -x -q -m gpu > gpurun_out/r2_1_bigparity.log 2>&1
echo "bigparity rc=$?" >> gpurun_out/r2_1_bigparity.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_1_bench_reblur.json 2> gpurun_out/r2_1_bench_reblur.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r2_1_bench_reference.json 2> gpurun_out/r2_1_bench_reference.err
timeout 300 python bench.py --denoiser relax --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_1_bench_relax.json 2> gpurun_out/r2_1_bench_relax.err
timeout 300 python bench.py --denoiser sigma --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_1_bench_sigma.json 2> gpurun_out/r2_1_bench_sigma.err
timeout 1500 python -m pytest tests -x -q -m gpu --deselect tests/test_parity_at_baseline_sizes_gpu.py > gpurun_out/r2_1_gputests.log 2>&1
echo "gputests rc=$?" >> gpurun_out/r2_1_gputests.log
tail -3 gpurun_out/r2_1_bigparity.log gpurun_out/r2_1_gputests.log
