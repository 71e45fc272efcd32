#!/usr/bin/env bash
# On the GPU box: short bench of the main build and of every prebuilt variant (tools/build_variant.py), per-pass microseconds.
#   tools/run_variants.sh [bench.py args]
run() {
  NRD_B200_LIB="$2" python bench.py --steps 24 --warmup 6 --no-cpu-baseline "${@:3}" 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1', 'ms/step', round(d['ms_per_step'],4), {k:round(v['avg_us'],1) for k,v in d['roofline']['passes'].items()})"
}
run main "" "$@"
for lib in nrd_sample_b200/variants/libnrd_b200_*.so; do
  n=$(basename "$lib" .so); run "${n#libnrd_b200_}" "$PWD/$lib" "$@"
done
run main-again "" "$@"
