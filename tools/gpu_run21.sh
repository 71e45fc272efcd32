#!/bin/bash
mkdir -p gpurun_out/ev11
O=gpurun_out/ev11
timeout 300 python -m pytest tests/test_relax_parity_gpu.py -q -m gpu -x > $O/relax_tests.log 2>&1; echo "rc=$?" >> $O/relax_tests.log
timeout 300 python -m pytest tests/test_reference_shaders_parity_gpu.py -q -m gpu -k "relax" >> $O/relax_tests.log 2>&1; echo "rc=$?" >> $O/relax_tests.log
grep -E "passed|failed|rc=|^E  " $O/relax_tests.log | cut -c1-300
bash tools/gpu_run20.sh "$@"
