#!/bin/bash
# a-trous kernels templated on enableRoughnessEdgeStopping: RELAX parity in one process, then a short bench if the budget allows
mkdir -p gpurun_out/ev16
O=gpurun_out/ev16
timeout 70 python -m pytest tests/test_relax_parity_gpu.py tests/test_relax_tma_gpu.py tests/test_reference_shaders_parity_gpu.py -k "relax" -q -m gpu -x > $O/tests.log 2>&1; echo "rc=$?" >> $O/tests.log
grep -E "passed|failed|rc=|^E  " $O/tests.log | tail -4 | cut -c1-250
timeout 40 python bench.py --denoiser relax --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_relax.json 2> $O/err
python - <<PY
import json
try:
    d=json.loads(open("$O/bench_relax.json").read().strip().splitlines()[-1])
    print("relax", round(d["ms_per_step"],4), {k:v["avg_us"] for k,v in d["roofline"]["passes"].items()})
except Exception as e: print("relax", "ERR", e)
PY
