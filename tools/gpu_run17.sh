#!/bin/bash
# RELAX parity + short bench ( per-pass microseconds ) of the build in the tree
mkdir -p gpurun_out/ev8
O=gpurun_out/ev8
timeout 600 python -m pytest tests/test_relax_parity_gpu.py tests/test_strips_sigma_relax_gpu.py -q -m gpu -x > $O/relax_tests.log 2>&1; echo "rc=$?" >> $O/relax_tests.log
timeout 600 python -m pytest tests/test_reference_shaders_parity_gpu.py -q -m gpu -k "relax" >> $O/relax_tests.log 2>&1; echo "rc=$?" >> $O/relax_tests.log
grep -E "passed|failed|rc=|^E  " $O/relax_tests.log | cut -c1-300
timeout 300 python bench.py --denoiser relax --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_relax.json 2> $O/err1
python - <<PY
import json
d=json.loads(open("$O/bench_relax.json").read().strip().splitlines()[-1])
print("relax", round(d["ms_per_step"],4), {k:v["avg_us"] for k,v in d["roofline"]["passes"].items()})
PY
