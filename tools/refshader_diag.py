"""Diagnostic: where do the oracle and the compiled reference shaders differ on this machine? python tools/refshader_diag.py [relax|reblur|sigma]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from nrd_sample_b200 import nrd_api as api, synth  # noqa: E402
from oracle import runner  # noqa: E402
from tests.test_oracle_vs_reference_shaders import DENOISERS, make_denoiser  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "relax"
w, h = 96, 64
os.system("grep -m1 'model name' /proc/cpuinfo; grep -m1 -o 'fma' /proc/cpuinfo | head -1; ldd --version | head -1")
den = make_denoiser(which, w, h)
fn = runner.ref_shaders().nrd_refshader_dispatch
snap = {}
shown = [0]


def before(i, d, keys, self):
    snap["t"] = [self.textures[k].clone() for k in keys]


def after(i, d, keys, self):
    ref = snap["t"]
    arr = (runner.OracleTexture * len(keys))(*[runner.tex_desc(t, self.formats[k]) for t, k in zip(ref, keys)])
    cb = C.create_string_buffer(d.constants, len(d.constants)) if d.constants else None
    assert fn(d.shader.encode(), cb, len(d.constants), arr, len(keys), d.grid[0], d.grid[1], 0) == 0
    for j, (b, k) in enumerate(zip(d.bindings, keys)):
        if b.descriptor != int(api.DescriptorType.STORAGE_TEXTURE) or torch.equal(ref[j], self.textures[k]) or shown[0] > 6:
            continue
        shown[0] += 1
        a, o = ref[j], self.textures[k]
        ne = (a != o)
        if ne.dim() == 3:
            ne = ne.any(-1)
        idx = ne.nonzero()[:4]
        print(f"frame {frame} {d.name} out{j} {api.Format(self.formats[k]).name}: {ne.float().mean().item():.2e} differ")
        for y, x in idx.tolist():
            print("   ", (x, y), "reference:", a[y, x].tolist() if a.dim() == 3 else a[y, x].item(), "oracle:", o[y, x].tolist() if o.dim() == 3 else o[y, x].item())


for frame in range(3):
    for k, v in DENOISERS[which][1](frame, w, h).items():
        den.set_user_texture(getattr(api.ResourceType, k), v)
    den.denoise(synth.common_settings(frame, w, h), before_dispatch=before, on_dispatch=after)
print("differences shown:", shown[0])
