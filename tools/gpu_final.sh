#!/bin/bash
# Final confirmation on ONE B200: the whole GPU suite, smoke, the four bench lines
mkdir -p gpurun_out/fin
O=gpurun_out/fin
timeout 2400 python -m pytest tests -q -m gpu > $O/gputests.log 2>&1; echo "gputests rc=$?" >> $O/gputests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_reblur.json 2> $O/bench_reblur.err
timeout 900 python bench.py --denoiser relax --steps 20 --warmup 5 > $O/bench_relax.json 2> $O/bench_relax.err
timeout 900 python bench.py --denoiser sigma --steps 40 --warmup 10 > $O/bench_sigma.json 2> $O/bench_sigma.err
timeout 900 python bench.py --denoiser reblur_sh --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_reblur_sh.json 2> $O/bench_reblur_sh.err
grep -E "passed|failed|rc=|^E  |^FAILED" $O/gputests.log $O/smoke.log | cut -c1-300 | tail -12
for f in reblur relax sigma reblur_sh; do python - <<PY
import json
try:
    d=json.loads(open("$O/bench_$f.json").read().strip().splitlines()[-1])
    print("$f", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["ms_per_step"],3), "graph", round(d["cuda_graph"]["ms_per_step"],4), "vs", d["vs_baseline"], d["clocks"]["samples"], {k:v["avg_us"] for k,v in d["roofline"]["passes"].items()})
except Exception as e: print("$f", "ERR", e)
PY
done
