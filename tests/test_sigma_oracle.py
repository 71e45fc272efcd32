"""CPU-side checks of the SIGMA_SHADOW oracle (oracle/sigma_passes.cpp) and of the synthetic shadow generator:
a regression fixture, behavioural properties (the pin against the reference's shaders is tests/test_oracle_vs_reference_shaders.py), and
BASELINE.json config 0: the single-pass SIGMA blur on a 512x512 tile as a CPU scalar run."""
import ctypes as C
import os
import time

import torch

from nrd_sample_b200 import nrd_api as api, synth
from oracle import runner

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sigma_96x64.pt")
R8 = api.Format.R8_UNORM
RT = api.ResourceType


def make(w, h):
    den = runner.OracleDenoiser(runner.default_host_library(), api.Denoiser.SIGMA_SHADOW, w, h)
    out = runner.alloc_texture(R8, w, h)
    den.set_user_texture(RT.OUT_SHADOW_TRANSLUCENCY, out)
    return den, out


def feed(den, frame):
    for k, v in frame.items():
        if not k.startswith("_"):
            den.set_user_texture(getattr(RT, k), v)


def test_generator_and_oracle_reproduce_golden_fixture():
    g = torch.load(GOLDEN)
    w, h = g["width"], g["height"]
    den, out = make(w, h)
    for f, stored in enumerate(g["inputs"]):
        fresh = synth.sigma_frame(f, w, h)
        for k in stored:
            assert torch.equal(fresh[k].view(torch.uint8), stored[k].view(torch.uint8)), f"generator drifted: frame {f} {k}"
        feed(den, {k: v.clone() for k, v in stored.items()})
        den.denoise(synth.common_settings(f, w, h))
        diff = (out.to(torch.int32) - g["outputs"][f].to(torch.int32)).abs()
        assert diff.max().item() <= 1 and (diff > 0).float().mean().item() < 1e-3, f"frame {f}: {diff.max().item()} LSB"


def test_dispatch_list_and_clears():
    den, _ = make(96, 64)
    feed(den, synth.sigma_frame(0, 96, 64))
    d0 = den.denoise(synth.common_settings(0, 96, 64))
    names = [d.shader.split("|")[0] for d in d0]
    assert names[-6:] == ["SIGMA_ClassifyTiles.cs.hlsl", "SIGMA_SmoothTiles.cs.hlsl", "SIGMA_Copy.cs.hlsl", "SIGMA_Blur.cs.hlsl", "SIGMA_Blur.cs.hlsl",
                          "SIGMA_TemporalStabilization.cs.hlsl"]
    assert all(n.startswith("Clear") for n in names[:-6]) and len(names) > 6
    feed(den, synth.sigma_frame(1, 96, 64))
    assert len(den.denoise(synth.common_settings(1, 96, 64))) == 6


def test_denoising_reduces_error_and_keeps_hard_regions():
    w, h, n = 160, 96, 8
    den, out = make(w, h)
    for f in range(n):
        fr = synth.sigma_frame(f, w, h, with_clean=(f == n - 1))
        feed(den, fr)
        den.denoise(synth.common_settings(f, w, h))
    m, clean = fr["_hit"], fr["_clean_visibility"]
    pen = fr["IN_PENUMBRA"].float()
    noisy = (pen >= synth.FP16_MAX).float()
    shadow = (out.float() / 255.0) ** 2   # SIGMA_BackEnd_UnpackShadow
    mse = lambda a, mask: ((a - clean) ** 2)[mask].mean().item()  # noqa: E731
    assert mse(shadow, m) < mse(noisy, m) / 6.0, (mse(noisy, m), mse(shadow, m))
    soft = m & (clean > 0.02) & (clean < 0.98)
    assert soft.float().mean().item() > 0.02, "the synthetic scene must contain penumbra"
    assert mse(shadow, soft) < mse(noisy, soft) / 8.0
    # pixels facing away from the light stay exactly black; the history length plane holds viewZ with 0..7 in the low bits
    away = m & (pen == 0.0)
    assert (out[away] == 0).all()
    hl = den.textures[(int(RT.PERMANENT_POOL), 0)]
    z = (hl & ~7).view(torch.float32)
    assert torch.allclose(z[m], (fr["IN_VIEWZ"].view(torch.int32)[m] & ~7).view(torch.float32))
    assert ((hl & 7)[m] >= 1).all()


def test_fully_lit_and_fully_shadowed_inputs_are_fixed_points():
    w, h = 96, 64
    for value, expect in ((synth.FP16_MAX, 255), (0.0, 0)):
        den, out = make(w, h)
        for f in range(3):
            fr = synth.sigma_frame(f, w, h)
            fr["IN_PENUMBRA"] = torch.full((h, w), value, dtype=torch.float16)
            feed(den, fr)
            den.denoise(synth.common_settings(f, w, h))
        hit = fr["IN_VIEWZ"] < 5e5
        assert (out[hit] == expect).all()


def test_config0_single_pass_blur_512_scalar():
    """BASELINE.json configs[0]: SIGMA_Shadow single-pass blur (SIGMA_Blur FIRST_PASS=1) on a 512x512 shadow + viewZ tile, one thread."""
    w = h = 512
    den, out = make(w, h)
    feed(den, synth.sigma_frame(0, w, h))
    blur = {}

    def grab(i, d, keys, self):
        if d.shader == "SIGMA_Blur.cs.hlsl|TRANSLUCENCY=0|FIRST_PASS=1":
            blur["d"], blur["keys"] = d, keys
            blur["inputs"] = [self.textures[k].clone() for k in keys]

    def grab_out(i, d, keys, self):
        if d.shader == "SIGMA_Blur.cs.hlsl|TRANSLUCENCY=0|FIRST_PASS=1":
            blur["outputs"] = [self.textures[k].clone() for k in keys]

    den.denoise(synth.common_settings(0, w, h), before_dispatch=grab, on_dispatch=grab_out)
    L = runner.lib()
    threads = L.nrd_oracle_max_threads()
    L.nrd_oracle_set_threads(1)
    try:
        tex = [t.clone() for t in blur["inputs"]]
        arr = (runner.OracleTexture * len(tex))(*[runner.tex_desc(t, den.formats[k]) for t, k in zip(tex, blur["keys"])])
        d = blur["d"]
        t0 = time.time()
        rc = L.nrd_oracle_dispatch(d.shader.encode(), C.create_string_buffer(d.constants, len(d.constants)), len(d.constants), arr, len(tex), d.grid[0], d.grid[1], 1)
        dt = time.time() - t0
    finally:
        L.nrd_oracle_set_threads(threads)
    assert rc == 0 and dt < 30.0
    # scalar replay == the OpenMP run of the same dispatch, bit for bit
    for got, want in zip(tex, blur["outputs"]):
        assert torch.equal(got.view(torch.uint8), want.view(torch.uint8))
    shadow, penumbra = tex[-1], tex[-2]
    hit = blur["inputs"][0] < 5e5
    pen_in = blur["inputs"][2].float()
    assert (shadow[hit & (pen_in == 0.0)] == 0).all(), "surfaces facing away from the light stay black"
    assert (penumbra.float()[hit] >= 0).all()
    soft = hit & (shadow > 0) & (shadow < 255)
    assert soft.float().mean().item() > 0.01, "the blur must produce intermediate shadow values in the penumbra"
    print(f"config0: SIGMA_Blur 512x512 single thread {dt * 1e3:.1f} ms = {w * h / dt / 1e6:.2f} Mpx/s")
