#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/tiled_check.py 3840 2160 4 10 > gpurun_out/r2_5_tiled_check_n$N.log 2>&1; tail -n 3 gpurun_out/r2_5_tiled_check_n$N.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_5_bench_n$N.json 2> gpurun_out/r2_5_bench_n$N.err
echo "rc=$?"; tail -c 1500 gpurun_out/r2_5_bench_n$N.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_5_bench_n$N.json").read().strip().splitlines()[-1])
    print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"].get("pcie_gbs_per_gpu"), "numa", d["e2e"].get("host_numa_node_rank0"))
    print("tiled", d.get("tiled_4k"))
except Exception as e:
    print("parse failed", e)
PY
