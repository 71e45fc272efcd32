#!/bin/bash
# parity of the main build ( unconditional tap loads in the gathering a-trous ), then main vs the HF_FETCH_ALL variant for the three chains
mkdir -p gpurun_out/ev10
O=gpurun_out/ev10
timeout 300 python -m pytest tests/test_relax_parity_gpu.py -q -m gpu -x > $O/relax_tests.log 2>&1; echo "rc=$?" >> $O/relax_tests.log
timeout 300 python -m pytest tests/test_reference_shaders_parity_gpu.py -q -m gpu -k "relax" >> $O/relax_tests.log 2>&1; echo "rc=$?" >> $O/relax_tests.log
grep -E "passed|failed|rc=|^E  " $O/relax_tests.log | cut -c1-300
run() {
  NRD_B200_LIB="$2" python bench.py --steps 24 --warmup 6 --no-cpu-baseline "${@:3}" 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$1', '${@:3}', 'ms/step', round(d['ms_per_step'],4), {k:round(v['avg_us'],1) for k,v in d['roofline']['passes'].items()})"
}
V=$PWD/nrd_sample_b200/variants/libnrd_b200_hfall.so
for den in reblur relax sigma; do
  run main "" --denoiser $den
  run hfall $V --denoiser $den
done
