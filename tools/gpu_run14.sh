#!/bin/bash
mkdir -p gpurun_out/ev4
O=gpurun_out/ev4
timeout 600 python bench.py --denoiser relax --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_relax_main.json 2> $O/err1
NRD_B200_LIB=$PWD/nrd_sample_b200/variants/libnrd_b200_plainhelpers.so timeout 600 python bench.py --denoiser relax --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_relax_plain.json 2> $O/err2
for f in main plain; do python - <<PY
import json
d=json.loads(open("$O/bench_relax_$f.json").read().strip().splitlines()[-1])
print("$f", round(d["ms_per_step"],4), {k:v["avg_us"] for k,v in d["roofline"]["passes"].items()})
PY
done
