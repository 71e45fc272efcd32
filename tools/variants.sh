#!/usr/bin/env bash
# A/B harness for the GPU box: rebuild with extra nvcc flags, run a short bench, print per-pass microseconds.
#   tools/variants.sh "-DSPATIAL_MIN_BLOCKS=3" "-DSPATIAL_MIN_BLOCKS=4"
for flags in "$@"; do
  NRD_B200_NVCC_EXTRA="$flags" python -m nrd_sample_b200.build --force > /dev/null 2>&1 || { echo "build failed for $flags"; continue; }
  grep -E "Used [0-9]+ registers" nrd_sample_b200/build/build.log | tr '\n' ' ' | sed 's/ptxas info    : Used //g; s/registers, used [0-9] barriers//g; s/bytes smem//g'
  echo
  python bench.py --steps 20 --warmup 6 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$flags', 'ms/step', round(d['ms_per_step'],3), {k:v['avg_us'] for k,v in d['roofline']['passes'].items()})"
done
