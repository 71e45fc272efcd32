#!/bin/bash
# N GPUs of one box ( gpurun --gpus N ): strips of every denoiser family against the whole-frame run ( bit exact ), then the bench line with tiled_4k
N=${1:-2}
mkdir -p gpurun_out/mg
O=gpurun_out/mg
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
for what in reblur sigma relax; do
  timeout 600 $TR --master-port 29511 tools/tiled_check.py 3840 2160 4 20 peer balanced $what > $O/tiled_check_${what}_${N}gpu.log 2>&1
  echo "$what rc=$?" >> $O/tiled_check_${what}_${N}gpu.log
  grep -E '^\{"check"|rc=|differ|Error' $O/tiled_check_${what}_${N}gpu.log | cut -c1-700
done
timeout 900 $TR --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_reblur_${N}gpu.json 2> $O/bench_reblur_${N}gpu.err
tail -c 1500 $O/bench_reblur_${N}gpu.json
tail -3 $O/bench_reblur_${N}gpu.err
