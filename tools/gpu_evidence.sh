#!/bin/bash
# Round-end evidence on ONE B200: full GPU suite, bench lines of the four workloads, the reference arm, ncu launch list + --set full captures, small-frame latency.
mkdir -p gpurun_out/ev
O=gpurun_out/ev
timeout 2400 python -m pytest tests -q -m gpu > $O/gputests.log 2>&1; echo "gputests rc=$?" >> $O/gputests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > $O/bench_reblur.json 2> $O/bench_reblur.err
timeout 900 python bench.py --denoiser relax --steps 20 --warmup 5 > $O/bench_relax.json 2> $O/bench_relax.err
timeout 900 python bench.py --denoiser sigma --steps 40 --warmup 10 > $O/bench_sigma.json 2> $O/bench_sigma.err
timeout 900 python bench.py --denoiser reblur_sh --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_reblur_sh.json 2> $O/bench_reblur_sh.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err
timeout 120 tools/frame_latency.bin > $O/frame_latency.json 2> $O/frame_latency.err
# the ncu launch list of the bench command ( per-launch times are cold-cache and serialised: shares, not absolutes )
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"reblur|clearKernel" -c 400 --csv --log-file $O/bench_launches_ncu.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:reblur -s 32 -c 16 -f -o /tmp/reblur_full python tools/profile_frame.py 2560 1440 6 > $O/ncu_reblur.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:relax -s 40 -c 20 -f -o /tmp/relax_full python tools/profile_frame.py 2560 1440 6 relax > $O/ncu_relax.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sigma -s 20 -c 10 -f -o /tmp/sigma_full python tools/profile_frame.py 2560 1440 6 sigma > $O/ncu_sigma.log 2>&1
# the .ncu-rep files stay on the box ( 50 MB each ): their summaries and the two per-launch tables travel
for f in reblur relax sigma; do
  python tools/ncu_summary.py /tmp/${f}_full.ncu-rep "${f}(Blur|TemporalAccumulation|Atrous)" > $O/${f}_1440p_ncu_summary.txt 2>&1
  python tools/ncu_to_json.py /tmp/${f}_full.ncu-rep $f 2560x1440 >> $O/ncu_to_json.log 2>&1
done
cp profiles/dram_traffic.json profiles/inst_counts.json $O/ 2>/dev/null
grep -E "passed|failed|rc=" $O/gputests.log $O/smoke.log | tail -6
for f in reblur relax sigma reblur_sh; do python - <<PY
import json
try:
    d=json.loads(open("$O/bench_$f.json").read().strip().splitlines()[-1])
    print("$f", round(d["ms_per_step"],4), "e2e", round(d["e2e"]["ms_per_step"],3), "graph", round(d["cuda_graph"]["ms_per_step"],4), "vs", d["vs_baseline"], {k:v["avg_us"] for k,v in d["roofline"]["passes"].items()})
except Exception as e: print("$f", "ERR", e)
PY
done
tail -c 600 $O/bench_reference_arm.json; cat $O/frame_latency.json | cut -c1-400; ls -la $O | head -30
