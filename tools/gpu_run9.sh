#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_cuda_graph_gpu.py -q -m gpu > gpurun_out/r2_9_graph.log 2>&1
echo "graph rc=$?" >> gpurun_out/r2_9_graph.log
timeout 120 tools/frame_latency.bin > gpurun_out/r2_9_frame_latency.json 2> gpurun_out/r2_9_frame_latency.err
timeout 1800 python -m pytest tests -q -m gpu -x --deselect tests/test_parity_at_baseline_sizes_gpu.py > gpurun_out/r2_9_gputests.log 2>&1
echo "gputests rc=$?" >> gpurun_out/r2_9_gputests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_9_bench_reblur.json 2> gpurun_out/r2_9_bench_reblur.err
grep -E "^E  .*Error|passed|failed|^FAILED" gpurun_out/r2_9_graph.log gpurun_out/r2_9_gputests.log | cut -c1-400 | head -40
cat gpurun_out/r2_9_frame_latency.json gpurun_out/r2_9_frame_latency.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_9_bench_reblur.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], {k:v["avg_us"] for k,v in d["roofline"]["passes"].items()}, d["e2e"]["ms_per_step"])
PY
