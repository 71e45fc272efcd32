"""Runs N frames of REBLUR_DIFFUSE_SPECULAR at WxH on cuda:0 — the workload ncu attaches to (see profiles/README.md).
  ncu --set full -k regex:reblur -s 63 -c 7 ... python tools/profile_frame.py 2560 1440 10"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from nrd_sample_b200 import executor as ex, nrd_api as api, synth  # noqa: E402

W, H, N = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
dev = "cuda:0"
F16 = api.Format.RGBA16_SFLOAT
FMT = {"IN_VIEWZ": api.Format.R32_SFLOAT, "IN_NORMAL_ROUGHNESS": api.Format.R10_G10_B10_A2_UNORM, "IN_MV": F16, "IN_DIFF_RADIANCE_HITDIST": F16, "IN_SPEC_RADIANCE_HITDIST": F16}
frames = [synth.reblur_frame(i, W, H, device=dev, period=4) for i in range(4)]
den = ex.CudaDenoiser(api.Denoiser.REBLUR_DIFFUSE_SPECULAR, W, H)
od, os_ = ex.alloc_texture(F16, W, H, dev), ex.alloc_texture(F16, W, H, dev)
den.set_user_texture(api.ResourceType.OUT_DIFF_RADIANCE_HITDIST, od, F16)
den.set_user_texture(api.ResourceType.OUT_SPEC_RADIANCE_HITDIST, os_, F16)
torch.cuda.synchronize()
for i in range(N):
    for k, v in frames[i % 4].items():
        den.set_user_texture(getattr(api.ResourceType, k), v, FMT[k])
    den.set_common_settings(synth.common_settings(i, W, H, period=4))
    den.denoise()
torch.cuda.synchronize()
print("done", ex.launch_count())
