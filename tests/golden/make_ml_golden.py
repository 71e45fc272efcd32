"""Records tests/golden/ml_vectors.json: inputs + bit patterns of the results of the REFERENCE MathLib functions
(oracle/_ref/libml_ref.so = External/NRIFramework/External/MathLib/ml.hlsli compiled as C++). Build-container only."""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import runner  # noqa: E402
from tests.test_oracle_math import GOLDEN, SPEC, call, make_inputs  # noqa: E402

runner.build()
ref = C.CDLL(os.path.join(os.path.dirname(runner.REF_LIB_PATH), "libml_ref.so"))
rng = np.random.default_rng(20260925)
out = {name: [[row, call(ref, "ml_", name, row)] for row in make_inputs(name, rng, 24)] for name in sorted(SPEC)}
json.dump(out, open(GOLDEN, "w"), separators=(",", ":"))
print("wrote", GOLDEN, os.path.getsize(GOLDEN), "bytes")
