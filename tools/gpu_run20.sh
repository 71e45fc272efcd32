#!/bin/bash
# main vs prebuilt variants, one bench line each: tools/gpu_run20.sh <denoiser>:<variant>,<variant>... ...
run() {
  NRD_B200_LIB="$2" python bench.py --steps 24 --warmup 6 --no-cpu-baseline --denoiser $3 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$3', '$1', 'ms/step', round(d['ms_per_step'],4), {k:round(v['avg_us'],1) for k,v in d['roofline']['passes'].items()})"
}
for spec in "$@"; do
  den=${spec%%:*}
  run main "" $den
  for v in $(echo ${spec#*:} | tr ',' ' '); do run $v $PWD/nrd_sample_b200/variants/libnrd_b200_$v.so $den; done
done
