#!/bin/bash
# Evidence after the RELAX changes of the second session: full GPU suite, smoke, RELAX bench line + ncu --set full capture, REBLUR / SIGMA bench lines ( kernels unchanged ) as a regression check
mkdir -p gpurun_out/ev12
O=gpurun_out/ev12
timeout 1200 python -m pytest tests -q -m gpu > $O/gputests.log 2>&1; echo "gputests rc=$?" >> $O/gputests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
timeout 600 python bench.py --denoiser relax --steps 20 --warmup 5 > $O/bench_relax.json 2> $O/bench_relax.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_reblur.json 2> $O/bench_reblur.err
timeout 300 python bench.py --denoiser sigma --steps 40 --warmup 10 --no-cpu-baseline > $O/bench_sigma.json 2> $O/bench_sigma.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:relax -s 40 -c 20 -f -o /tmp/relax_full python tools/profile_frame.py 2560 1440 6 relax > $O/ncu_relax.log 2>&1
python tools/ncu_summary.py /tmp/relax_full.ncu-rep "relax(TemporalAccumulation|Atrous|PrePass|HistoryClamping)" > $O/relax_1440p_ncu_summary.txt 2>&1
python tools/ncu_to_json.py /tmp/relax_full.ncu-rep relax 2560x1440 > $O/ncu_to_json.log 2>&1
cp profiles/dram_traffic.json profiles/inst_counts.json $O/ 2>/dev/null
grep -E "passed|failed|rc=" $O/gputests.log $O/smoke.log | tail -6
for f in reblur relax sigma; do python - <<PY
import json
try:
    d=json.loads(open("$O/bench_$f.json").read().strip().splitlines()[-1])
    print("$f", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["ms_per_step"],3), "graph", round(d["cuda_graph"]["ms_per_step"],4), "vs", d["vs_baseline"], {k:v["avg_us"] for k,v in d["roofline"]["passes"].items()})
except Exception as e: print("$f", "ERR", e)
PY
done
