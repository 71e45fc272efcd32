#!/bin/bash
# TMA staging of the raw planes in the tiled a-trous kernels: parity, then A/B bench ( NRD_B200_RELAX_TMA=0 keeps the LDG staging )
mkdir -p gpurun_out/ev9
O=gpurun_out/ev9
timeout 300 python -m pytest tests/test_relax_parity_gpu.py tests/test_strips_sigma_relax_gpu.py -q -m gpu -x > $O/relax_tests.log 2>&1; echo "rc=$?" >> $O/relax_tests.log
timeout 300 python -m pytest tests/test_reference_shaders_parity_gpu.py tests/test_cuda_graph_gpu.py -q -m gpu -k "relax" >> $O/relax_tests.log 2>&1; echo "rc=$?" >> $O/relax_tests.log
grep -E "passed|failed|rc=|^E  " $O/relax_tests.log | cut -c1-300
timeout 200 python bench.py --denoiser relax --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_relax_tma.json 2> $O/err1
NRD_B200_RELAX_TMA=0 timeout 200 python bench.py --denoiser relax --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_relax_ldg.json 2> $O/err2
for f in tma ldg; do python - <<PY
import json
d=json.loads(open("$O/bench_relax_$f.json").read().strip().splitlines()[-1])
print("$f", round(d["ms_per_step"],4), {k:v["avg_us"] for k,v in d["roofline"]["passes"].items()})
PY
done
tail -3 $O/err1
