#!/bin/bash
# Programmatic dependent launch: full GPU suite + smoke on the build, then A/B bench lines ( NRD_B200_PDL=0 = plain stream serialization )
mkdir -p gpurun_out/ev13
O=gpurun_out/ev13
NRD_B200_PDL=1 timeout 900 python -m pytest tests -q -m gpu -x > $O/gputests.log 2>&1; echo "gputests rc=$?" >> $O/gputests.log
NRD_B200_PDL=1 timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
grep -E "passed|failed|rc=|^E  " $O/gputests.log $O/smoke.log | tail -8 | cut -c1-250
for den in reblur sigma relax; do
  NRD_B200_PDL=1 timeout 200 python bench.py --denoiser $den --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_${den}_pdl.json 2> $O/err_${den}_pdl
  [ $den = relax ] || NRD_B200_PDL=0 timeout 200 python bench.py --denoiser $den --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_${den}_plain.json 2> $O/err_${den}_plain
done
for f in reblur_pdl reblur_plain sigma_pdl sigma_plain relax_pdl; do python - <<PY
import json
try:
    d=json.loads(open("$O/bench_$f.json").read().strip().splitlines()[-1])
    print("$f", round(d["ms_per_step"],4), "graph", round(d["cuda_graph"]["ms_per_step"],4), "events-leg", round(d.get("ms_per_step_with_per_pass_events",0),4), "e2e", round(d["e2e"]["ms_per_step"],3))
except Exception as e: print("$f", "ERR", e)
PY
done
