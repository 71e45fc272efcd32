"""Which fields of REBLUR data2 differ between CUDA and oracle (strict mode)?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nrd_sample_b200 import executor as ex, nrd_api as api, synth
from oracle import runner
from tests.util import decode
W, H = 208, 120
orc = runner.OracleDenoiser(runner.default_host_library(), api.Denoiser.REBLUR_DIFFUSE_SPECULAR, W, H, robust_mirror_test=True)
orc.set_user_texture(api.ResourceType.OUT_DIFF_RADIANCE_HITDIST, runner.alloc_texture(api.Format.RGBA16_SFLOAT, W, H))
orc.set_user_texture(api.ResourceType.OUT_SPEC_RADIANCE_HITDIST, runner.alloc_texture(api.Format.RGBA16_SFLOAT, W, H))
snap = {}
def before(i, d, keys, den): snap['t'] = [den.textures[k].clone() for k in keys]
def after(i, d, keys, den):
    if 'Temporal accumulation' not in d.name: return
    gpu = [t.to('cuda:0') for t in snap['t']]
    ex.dispatch(d.shader, d.constants, [ex.texture_of(g, den.formats[k]) for g, k in zip(gpu, keys)], flags=5)
    torch.cuda.synchronize()
    a, b = decode(gpu[24], api.Format.R32_UINT), decode(den.textures[keys[24]], api.Format.R32_UINT)
    names = ['occlusion bits', 'virtualHistoryAmount', 'allowCatRom', 'curvature']
    for c in range(4):
        diff = (a[..., c] - b[..., c]).abs()
        print(f"  {names[c]:22s} differing px {int((diff > 0).sum())}  max {diff.max().item():.4g}  ref absmax {b[..., c].abs().max().item():.4g}")
    cd = (a[..., 3] - b[..., 3]).abs(); idx = cd.argmax(); y, x = divmod(int(idx), W)
    print('  worst curvature at', x, y, a[y, x, 3].item(), b[y, x, 3].item())
for f in range(4):
    for k, v in synth.reblur_frame(f, W, H).items(): orc.set_user_texture(getattr(api.ResourceType, k), v)
    print('frame', f); orc.denoise(synth.common_settings(f, W, H), before_dispatch=before, on_dispatch=after)
