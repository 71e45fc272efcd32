"""Rates of the "mirrored" branch of the spatial taps ( REBLUR_Common_SpatialFilter.hlsli:198 ) per pass and lobe, reference shaders vs CUDA, on one GPU:
  python tools/mirror_rates.py [W H [FRAMES]]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from nrd_sample_b200 import executor as ex, nrd_api as api, synth  # noqa: E402
from oracle import runner  # noqa: E402

W = int(sys.argv[1]) if len(sys.argv) > 1 else 960
H = int(sys.argv[2]) if len(sys.argv) > 2 else 540
N = int(sys.argv[3]) if len(sys.argv) > 3 else 3
RT, F16 = api.ResourceType, api.Format.RGBA16_SFLOAT
ref = runner.OracleDenoiser(runner.default_host_library(), api.Denoiser.REBLUR_DIFFUSE_SPECULAR, W, H, engine="reference")
cud = ex.CudaDenoiser(api.Denoiser.REBLUR_DIFFUSE_SPECULAR, W, H, flags=ex.FLAG_QUAD_INTRINSICS | ex.FLAG_PROBE_MIRROR)
outs = {}
for o in ("OUT_DIFF_RADIANCE_HITDIST", "OUT_SPEC_RADIANCE_HITDIST"):
    ref.set_user_texture(getattr(RT, o), runner.alloc_texture(F16, W, H), F16)
    outs[o] = ex.alloc_texture(F16, W, H, "cuda:0")
    cud.set_user_texture(getattr(RT, o), outs[o], F16)
runner.ref_mirror_probe(reset=True)
ex.mirror_probe(reset=True)
keep = {}
for f in range(N):
    for k, v in synth.reblur_frame(f, W, H).items():
        ref.set_user_texture(getattr(RT, k), v)
        keep[k] = v.to("cuda:0")
        cud.set_user_texture(getattr(RT, k), keep[k], runner.USER_FORMATS[getattr(RT, k)])
    cs = synth.common_settings(f, W, H)
    ref.denoise(cs)
    cud.set_common_settings(cs)
    cud.denoise()
torch.cuda.synchronize()
a, b = runner.ref_mirror_probe_detail(), ex.mirror_probe_detail()
for k in a:
    print(f"{k[0]:10s} {k[1]:5s} taps {a[k][0]:11d} / {b[k][0]:11d}   mirrored: reference {a[k][1] / max(a[k][0], 1):.4f}   cuda {b[k][1] / max(b[k][0], 1):.4f}")
