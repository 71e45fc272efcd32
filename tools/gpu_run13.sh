#!/bin/bash
mkdir -p gpurun_out/ev5
O=gpurun_out/ev5
timeout 1200 python -m pytest tests/test_relax_parity_gpu.py tests/test_strips_sigma_relax_gpu.py -q -m gpu > $O/relax_tests.log 2>&1; echo "rc=$?" >> $O/relax_tests.log
timeout 1200 python -m pytest tests/test_reference_shaders_parity_gpu.py -q -m gpu -k "relax" >> $O/relax_tests.log 2>&1; echo "rc=$?" >> $O/relax_tests.log
timeout 600 python bench.py --denoiser relax --steps 20 --warmup 5 > $O/bench_relax.json 2> $O/bench_relax.err
grep -E "passed|failed|rc=|^E  " $O/relax_tests.log | cut -c1-300
python - <<PY
import json
d=json.loads(open("$O/bench_relax.json").read().strip().splitlines()[-1])
print("relax", round(d["ms_per_step"],4), "vs", d["vs_baseline"], d["roofline"]["chain"], {k:(v["avg_us"],v.get("issue_frac")) for k,v in d["roofline"]["passes"].items()})
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:relax -s 40 -c 20 -f -o /tmp/relax_full python tools/profile_frame.py 2560 1440 6 relax > $O/ncu_relax.log 2>&1
python tools/ncu_summary.py /tmp/relax_full.ncu-rep "relax(TemporalAccumulation|Atrous|PrePass)" > $O/relax_1440p_ncu_summary.txt 2>&1
python tools/ncu_to_json.py /tmp/relax_full.ncu-rep relax 2560x1440 > $O/ncu_to_json.log 2>&1
cp profiles/dram_traffic.json profiles/inst_counts.json $O/
