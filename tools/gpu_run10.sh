#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sigma_parity_gpu.py tests/test_cuda_graph_gpu.py -q -m gpu > gpurun_out/r2_10_sigma.log 2>&1
echo "sigma rc=$?" >> gpurun_out/r2_10_sigma.log
timeout 900 python -m pytest tests/test_reference_shaders_parity_gpu.py -q -m gpu -k "sigma" >> gpurun_out/r2_10_sigma.log 2>&1
echo "sigma refshader rc=$?" >> gpurun_out/r2_10_sigma.log
timeout 600 python bench.py --denoiser sigma --steps 40 --warmup 10 --no-cpu-baseline > gpurun_out/r2_10_bench_sigma.json 2> gpurun_out/r2_10_bench_sigma.err
grep -E "^E  .*Error|passed|failed|^FAILED|rc=" gpurun_out/r2_10_sigma.log | cut -c1-400 | head -40
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_10_bench_sigma.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["vs_baseline"], {k:v["avg_us"] for k,v in d["roofline"]["passes"].items()}, d["e2e"]["ms_per_step"])
PY
