import sys, time, torch, ctypes as C
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nrd_sample_b200 import nrd_api as api, synth, executor as ex
from oracle import runner
from tests.util import compare
W, H, N = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
ROBUST = len(sys.argv) > 4 and sys.argv[4] == 'robust'
GFLAGS = ex.FLAG_QUAD_INTRINSICS | (ex.FLAG_ROBUST_MIRROR_TEST if ROBUST else 0)
dev = 'cuda:0'
host = runner.default_host_library()
orc = runner.OracleDenoiser(host, api.Denoiser.REBLUR_DIFFUSE_SPECULAR, W, H, robust_mirror_test=ROBUST)
o_d = runner.alloc_texture(api.Format.RGBA16_SFLOAT, W, H); o_s = runner.alloc_texture(api.Format.RGBA16_SFLOAT, W, H)
orc.set_user_texture(api.ResourceType.OUT_DIFF_RADIANCE_HITDIST, o_d); orc.set_user_texture(api.ResourceType.OUT_SPEC_RADIANCE_HITDIST, o_s)
worst = {}
def before(i, d, keys, den):
    den._snap = [den.textures[k].clone() for k in keys]
def after(i, d, keys, den):
    # replay this dispatch on the GPU from the oracle's pre-dispatch textures
    gpu = [t.to(dev) for t in den._snap]
    texs = [ex.texture_of(g, den.formats[k]) for g, k in zip(gpu, keys)]
    ex.dispatch(d.shader, d.constants, texs, flags=GFLAGS)
    torch.cuda.synchronize()
    for j, (b, k) in enumerate(zip(d.bindings, keys)):
        if b.descriptor != 1: continue
        r = compare(gpu[j], den.textures[k], den.formats[k])
        key = (d.name.split(' - ')[-1], j, api.Format(den.formats[k]).name)
        w = worst.get(key)
        if d.name.startswith('Clear'): continue
        if w is None or r['frac_bad'] > w['frac_bad']: worst[key] = r
for f in range(N):
    fr = synth.reblur_frame(f, W, H)
    for k, v in fr.items(): orc.set_user_texture(getattr(api.ResourceType, k), v)
    orc.denoise(synth.common_settings(f, W, H), before_dispatch=before, on_dispatch=after)
for k, r in worst.items():
    print(f"{k[0]:28s} binding {k[1]:2d} {k[2]:22s} frac_bad {r['frac_bad']:.2e} max_abs {r['max_abs']:.3e} psnr {r['psnr']:.1f}")
# closed loop
cud = ex.CudaDenoiser(api.Denoiser.REBLUR_DIFFUSE_SPECULAR, W, H, flags=GFLAGS)
orc2 = runner.OracleDenoiser(host, api.Denoiser.REBLUR_DIFFUSE_SPECULAR, W, H, robust_mirror_test=ROBUST)
o_d2 = runner.alloc_texture(api.Format.RGBA16_SFLOAT, W, H); o_s2 = runner.alloc_texture(api.Format.RGBA16_SFLOAT, W, H)
orc2.set_user_texture(api.ResourceType.OUT_DIFF_RADIANCE_HITDIST, o_d2); orc2.set_user_texture(api.ResourceType.OUT_SPEC_RADIANCE_HITDIST, o_s2)
g_d = ex.alloc_texture(api.Format.RGBA16_SFLOAT, W, H, dev); g_s = ex.alloc_texture(api.Format.RGBA16_SFLOAT, W, H, dev)
cud.set_user_texture(api.ResourceType.OUT_DIFF_RADIANCE_HITDIST, g_d, api.Format.RGBA16_SFLOAT); cud.set_user_texture(api.ResourceType.OUT_SPEC_RADIANCE_HITDIST, g_s, api.Format.RGBA16_SFLOAT)
for f in range(max(N, 12)):
    fr = synth.reblur_frame(f, W, H)
    gfr = {k: v.to(dev) for k, v in fr.items()}
    for k, v in fr.items():
        orc2.set_user_texture(getattr(api.ResourceType, k), v)
        cud.set_user_texture(getattr(api.ResourceType, k), gfr[k], runner.USER_FORMATS[getattr(api.ResourceType, k)])
    cs = synth.common_settings(f, W, H)
    orc2.denoise(cs); cud.set_common_settings(cs); cud.denoise(); torch.cuda.synchronize()
    rd = compare(g_d, o_d2, api.Format.RGBA16_SFLOAT); rs = compare(g_s, o_s2, api.Format.RGBA16_SFLOAT)
    print(f"closed loop frame {f}: diff psnr {rd['psnr']:.1f} bad {rd['frac_bad']:.2e} | spec psnr {rs['psnr']:.1f} bad {rs['frac_bad']:.2e}")
print('launches', ex.launch_count())
