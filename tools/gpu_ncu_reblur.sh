#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:reblur -s 72 -c 8 -f -o gpurun_out/reblur_r2_a python tools/profile_frame.py 2560 1440 10 > gpurun_out/r2_ncu_reblur.log 2>&1
tail -n 12 gpurun_out/r2_ncu_reblur.log
