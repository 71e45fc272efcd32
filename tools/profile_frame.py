"""Runs N frames of one denoiser at WxH on cuda:0 — the workload ncu attaches to (see profiles/README.md) — and prints
the per-pass CUDA-event times of the last N - 4 frames.
  ncu --set full -k regex:reblur -s 63 -c 7 ... python tools/profile_frame.py 2560 1440 10 [reblur|sigma|relax]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from nrd_sample_b200 import executor as ex, nrd_api as api, synth  # noqa: E402

W, H, N = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
WHAT = sys.argv[4] if len(sys.argv) > 4 else "reblur"
dev = "cuda:0"
F16 = api.Format.RGBA16_SFLOAT
RT = api.ResourceType
FMT = {"IN_VIEWZ": api.Format.R32_SFLOAT, "IN_NORMAL_ROUGHNESS": api.Format.R10_G10_B10_A2_UNORM, "IN_MV": F16, "IN_DIFF_RADIANCE_HITDIST": F16, "IN_SPEC_RADIANCE_HITDIST": F16,
       "IN_PENUMBRA": api.Format.R16_SFLOAT}
if WHAT == "relax":
    frames = [synth.relax_frame(i, W, H, device=dev, period=4) for i in range(4)]
    den = ex.CudaDenoiser(api.Denoiser.RELAX_DIFFUSE_SPECULAR_SH, W, H)
    outs = [ex.alloc_texture(F16, W, H, dev) for _ in range(4)]
    for rt, t in zip((RT.OUT_DIFF_SH0, RT.OUT_DIFF_SH1, RT.OUT_SPEC_SH0, RT.OUT_SPEC_SH1), outs):
        den.set_user_texture(rt, t, F16)
elif WHAT == "sigma":
    frames = [synth.sigma_frame(i, W, H, device=dev, period=4) for i in range(4)]
    den = ex.CudaDenoiser(api.Denoiser.SIGMA_SHADOW, W, H)
    outs = [ex.alloc_texture(api.Format.R8_UNORM, W, H, dev)]
    den.set_user_texture(RT.OUT_SHADOW_TRANSLUCENCY, outs[0], api.Format.R8_UNORM)
else:
    frames = [synth.reblur_frame(i, W, H, device=dev, period=4) for i in range(4)]
    den = ex.CudaDenoiser(api.Denoiser.REBLUR_DIFFUSE_SPECULAR, W, H)
    outs = [ex.alloc_texture(F16, W, H, dev), ex.alloc_texture(F16, W, H, dev)]
    den.set_user_texture(RT.OUT_DIFF_RADIANCE_HITDIST, outs[0], F16)
    den.set_user_texture(RT.OUT_SPEC_RADIANCE_HITDIST, outs[1], F16)
torch.cuda.synchronize()
for i in range(N):
    if i == 4:
        torch.cuda.synchronize()
        den.set_profiling(True)
    for k, v in frames[i % 4].items():
        den.set_user_texture(getattr(RT, k), v, FMT.get(k, F16))
    den.set_common_settings(synth.common_settings(i, W, H, period=4))
    den.denoise()
torch.cuda.synchronize()
prof = den.profile() if N > 4 else {}
total = 0.0
for name, (tot, cnt) in prof.items():
    if cnt:
        print(f"{name:44s} {tot / cnt * 1e3:9.2f} us  x{cnt}")
        total += tot / (N - 4)   # per frame: the a-trous pass runs several times per frame
if prof:
    print(f"frame total {total * 1e3:.1f} us = {W * H / total / 1e3:.0f} Mpx/s")
print("done", ex.launch_count())
