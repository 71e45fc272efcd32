#!/bin/bash
mkdir -p gpurun_out/ev2
O=gpurun_out/ev2
timeout 900 python -m pytest tests/test_sigma_parity_gpu.py tests/test_strips_sigma_relax_gpu.py tests/test_cuda_graph_gpu.py -q -m gpu > $O/sigma_tests.log 2>&1; echo "rc=$?" >> $O/sigma_tests.log
timeout 900 python -m pytest tests/test_reference_shaders_parity_gpu.py -q -m gpu -k sigma >> $O/sigma_tests.log 2>&1; echo "rc=$?" >> $O/sigma_tests.log
timeout 600 python bench.py --denoiser sigma --steps 40 --warmup 10 > $O/bench_sigma.json 2> $O/bench_sigma.err
timeout 600 python bench.py --steps 20 --warmup 5 > $O/bench_reblur.json 2> $O/bench_reblur.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"reblur|clearKernel" -c 400 --csv --log-file $O/bench_launches_ncu.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/bench_under_ncu.log 2>&1
grep -E "passed|failed|rc=|^E " $O/sigma_tests.log | cut -c1-300
for f in sigma reblur; do python - <<PY
import json
d=json.loads(open("$O/bench_$f.json").read().strip().splitlines()[-1])
print("$f", round(d["ms_per_step"],4), "vs", d["vs_baseline"], d["clocks"], d["roofline"]["chain"], {k:(v["avg_us"],v.get("issue_frac")) for k,v in d["roofline"]["passes"].items()})
PY
done
grep -c reblur $O/bench_launches_ncu.csv
