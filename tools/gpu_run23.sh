#!/bin/bash
# last check of the round: programmatic dependent launch on by default for stand-alone contexts
mkdir -p gpurun_out/ev14
O=gpurun_out/ev14
timeout 150 python -m pytest tests/test_cuda_graph_gpu.py tests/test_strips_sigma_relax_gpu.py tests/test_sigma_parity_gpu.py tests/test_relax_tma_gpu.py -q -m gpu -x > $O/tests.log 2>&1; echo "rc=$?" >> $O/tests.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
grep -E "passed|failed|rc=|^E  " $O/tests.log $O/smoke.log | tail -6 | cut -c1-250
for den in reblur sigma; do
  timeout 100 python bench.py --denoiser $den --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_${den}.json 2> $O/err_${den}
  python - <<PY
import json
try:
    d=json.loads(open("$O/bench_$den.json").read().strip().splitlines()[-1])
    print("$den", round(d["ms_per_step"],4), "graph", round(d["cuda_graph"]["ms_per_step"],4), "e2e", round(d["e2e"]["ms_per_step"],3), d["clocks"])
except Exception as e: print("$den", "ERR", e)
PY
done
