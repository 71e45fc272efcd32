#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_reference_shaders_parity_gpu.py -q -m gpu -k "validation" > gpurun_out/r2_8_validation.log 2>&1
echo "validation rc=$?" >> gpurun_out/r2_8_validation.log
timeout 1800 python -m pytest tests -q -m gpu --deselect tests/test_parity_at_baseline_sizes_gpu.py -k "not validation" > gpurun_out/r2_8_gputests.log 2>&1
echo "gputests rc=$?" >> gpurun_out/r2_8_gputests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_8_bench_reblur.json 2> gpurun_out/r2_8_bench_reblur.err
grep -E "^E  .*Error|passed|failed|^FAILED" gpurun_out/r2_8_validation.log gpurun_out/r2_8_gputests.log | cut -c1-400 | head -40
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_8_bench_reblur.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], {k:v["avg_us"] for k,v in d["roofline"]["passes"].items()}, d["e2e"]["ms_per_step"])
PY
