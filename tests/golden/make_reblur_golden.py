"""Writes tests/golden/reblur_96x64.pt: 4 frames of seeded synthetic REBLUR inputs (storage formats) and the oracle's
OUT_DIFF / OUT_SPEC_RADIANCE_HITDIST after each frame, in faithful and strict ("robust mirror test") modes.
PARITY UNPINNED: these vectors come from OUR oracle — the reference ships no images to compare with."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from nrd_sample_b200 import nrd_api as api, synth  # noqa: E402
from oracle import runner  # noqa: E402

W, H, N = 96, 64, 4
PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reblur_96x64.pt")


def run(robust):
    den = runner.OracleDenoiser(runner.default_host_library(), api.Denoiser.REBLUR_DIFFUSE_SPECULAR, W, H, robust_mirror_test=robust)
    od, os_ = runner.alloc_texture(api.Format.RGBA16_SFLOAT, W, H), runner.alloc_texture(api.Format.RGBA16_SFLOAT, W, H)
    den.set_user_texture(api.ResourceType.OUT_DIFF_RADIANCE_HITDIST, od)
    den.set_user_texture(api.ResourceType.OUT_SPEC_RADIANCE_HITDIST, os_)
    outs, ins = [], []
    for f in range(N):
        fr = synth.reblur_frame(f, W, H)
        ins.append({k: v.clone() for k, v in fr.items()})  # frame 0 clears IN_MV in place (it is bound read-write by TS)
        for k, v in fr.items():
            den.set_user_texture(getattr(api.ResourceType, k), v)
        den.denoise(synth.common_settings(f, W, H))
        outs.append((od.clone(), os_.clone()))
    return ins, outs


if __name__ == "__main__":
    ins, faithful = run(False)
    _, strict = run(True)
    torch.save({"width": W, "height": H, "inputs": ins, "faithful": faithful, "strict": strict}, PATH)
    print("wrote", PATH, os.path.getsize(PATH), "bytes")
