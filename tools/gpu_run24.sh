#!/bin/bash
# the bench line of the round-end build, default command ( what the driver runs )
mkdir -p gpurun_out/ev15
timeout 110 python bench.py --steps 20 --warmup 5 > gpurun_out/ev15/bench_reblur.json 2> gpurun_out/ev15/bench_reblur.err
tail -c 400 gpurun_out/ev15/bench_reblur.json
