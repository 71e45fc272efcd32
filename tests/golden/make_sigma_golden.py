"""Writes tests/golden/sigma_96x64.pt: 5 frames of seeded synthetic SIGMA_SHADOW inputs (storage formats) and the
oracle's OUT_SHADOW_TRANSLUCENCY after each frame.
PARITY UNPINNED: these vectors come from OUR oracle — the reference ships no images to compare with."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from nrd_sample_b200 import nrd_api as api, synth  # noqa: E402
from oracle import runner  # noqa: E402

W, H, N = 96, 64, 5
PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "sigma_96x64.pt")

if __name__ == "__main__":
    den = runner.OracleDenoiser(runner.default_host_library(), api.Denoiser.SIGMA_SHADOW, W, H)
    out = runner.alloc_texture(api.Format.R8_UNORM, W, H)
    den.set_user_texture(api.ResourceType.OUT_SHADOW_TRANSLUCENCY, out)
    ins, outs = [], []
    for f in range(N):
        fr = synth.sigma_frame(f, W, H)
        ins.append({k: v.clone() for k, v in fr.items()})
        for k, v in fr.items():
            den.set_user_texture(getattr(api.ResourceType, k), v)
        den.denoise(synth.common_settings(f, W, H))
        outs.append(out.clone())
    torch.save({"width": W, "height": H, "inputs": ins, "outputs": outs}, PATH)
    print("wrote", PATH, os.path.getsize(PATH), "bytes")
