"""Writes tests/golden/refshaders_{reblur,sigma,relax}_80x48.pt: the outputs of 4 frames denoised with the REFERENCE'S SHADERS as the engine
(oracle/_ref/libnrd_refshaders.so, compiled from /root/reference by oracle/ref_build_shaders.py; dispatch stream from the host library).
Run in the build container (the reference tree must be mounted): python tests/golden/make_refshader_golden.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from nrd_sample_b200 import nrd_api as api, synth  # noqa: E402
from oracle import runner  # noqa: E402
from tests.test_oracle_vs_reference_shaders import DENOISERS, make_denoiser  # noqa: E402

W, H, FRAMES = 80, 48, 4
assert runner.ref_shaders() is not None, "build oracle/_ref/libnrd_refshaders.so first"
for which in DENOISERS:
    den = make_denoiser(which, W, H, engine="reference")
    for frame in range(FRAMES):
        for k, v in DENOISERS[which][1](frame, W, H).items():
            den.set_user_texture(getattr(api.ResourceType, k), v)
        den.denoise(synth.common_settings(frame, W, H))
    outs = {o: den.textures[(int(getattr(api.ResourceType, o)), 0)].clone() for o in DENOISERS[which][2]}
    path = os.path.join(ROOT, "tests", "golden", f"refshaders_{which}_{W}x{H}.pt")
    torch.save({"width": W, "height": H, "frames": FRAMES, "engine": "reference shaders (libnrd_refshaders.so)", "outputs": outs}, path)
    print("wrote", path, os.path.getsize(path), "bytes")
