"""Writes tests/golden/relax_96x64.pt: 4 frames of seeded synthetic RELAX_DIFFUSE_SPECULAR_SH inputs (storage formats) and the
oracle's four OUT_*_SH* textures after each frame.
PARITY UNPINNED: these vectors come from OUR oracle — the reference ships no images to compare with."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from nrd_sample_b200 import nrd_api as api, synth  # noqa: E402
from oracle import runner  # noqa: E402

W, H, N = 96, 64, 4
PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "relax_96x64.pt")
OUTS = ("OUT_DIFF_SH0", "OUT_DIFF_SH1", "OUT_SPEC_SH0", "OUT_SPEC_SH1")

if __name__ == "__main__":
    den = runner.OracleDenoiser(runner.default_host_library(), api.Denoiser.RELAX_DIFFUSE_SPECULAR_SH, W, H)
    outs = {k: runner.alloc_texture(api.Format.RGBA16_SFLOAT, W, H) for k in OUTS}
    for k, v in outs.items():
        den.set_user_texture(getattr(api.ResourceType, k), v)
    ins, res = [], []
    for f in range(N):
        fr = synth.relax_frame(f, W, H)
        ins.append({k: v.clone() for k, v in fr.items()})
        for k, v in fr.items():
            den.set_user_texture(getattr(api.ResourceType, k), v)
        den.denoise(synth.common_settings(f, W, H))
        res.append({k: v.clone() for k, v in outs.items()})
    torch.save({"width": W, "height": H, "inputs": ins, "outputs": res}, PATH)
    print("wrote", PATH, os.path.getsize(PATH), "bytes")
