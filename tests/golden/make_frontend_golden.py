"""Writes tests/golden/frontend_probe.pt: the outputs of the REFERENCE's NRD.hlsli (compiled as C++, oracle/_ref/libnrd_refshaders.so) for the
verification sequence of include/nrd_frontend.cuh on 2048 seeded columns. Run in the build container (needs /root/reference):
    python tests/golden/make_frontend_golden.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests.test_frontend_codecs import GOLDEN, make_inputs, run_reference  # noqa: E402

n, seed = 2048, 4321
torch.save({"n": n, "seed": seed, "outputs": [t.clone() for t in run_reference(make_inputs(n, seed))]}, GOLDEN)
print("wrote", GOLDEN, os.path.getsize(GOLDEN), "bytes")
