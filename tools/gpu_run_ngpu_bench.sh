#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out/mg
O=gpurun_out/mg
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_reblur_${N}gpu.json 2> $O/bench_reblur_${N}gpu.err
tail -c 2500 $O/bench_reblur_${N}gpu.json
tail -3 $O/bench_reblur_${N}gpu.err
