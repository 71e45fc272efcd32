#!/bin/bash
# source-level captures ( --import-source on ) of the temporal-accumulation kernels and one gathering a-trous launch
mkdir -p gpurun_out/ev7
O=gpurun_out/ev7
timeout 600 ncu --set full --clock-control none --import-source on -k regex:relaxTemporalAccumulation -s 5 -c 1 -f -o $O/relax_ta python tools/profile_frame.py 2560 1440 8 relax > $O/ncu_relax_ta.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:reblurTemporalAccumulation -s 5 -c 1 -f -o $O/reblur_ta python tools/profile_frame.py 2560 1440 8 > $O/ncu_reblur_ta.log 2>&1
ls -la $O
