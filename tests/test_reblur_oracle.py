"""CPU-side checks of the REBLUR oracle (oracle/reblur_passes.cpp) and of the synthetic input generator.
The oracle is pinned three ways (none of them reference pixels — none exist): (1) its MathLib helpers against the
reference MathLib (test_oracle_math.py), (2) its dispatch stream / constants against the reference host library
(test_dispatch_stream.py), (3) behavioural properties and a regression fixture here."""
import os

import pytest
import torch

from nrd_sample_b200 import nrd_api as api, synth
from oracle import runner
from tests.util import compare

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reblur_96x64.pt")
F16 = api.Format.RGBA16_SFLOAT


def make(w, h, **kw):
    den = runner.OracleDenoiser(runner.default_host_library(), api.Denoiser.REBLUR_DIFFUSE_SPECULAR, w, h, **kw)
    od, os_ = runner.alloc_texture(F16, w, h), runner.alloc_texture(F16, w, h)
    den.set_user_texture(api.ResourceType.OUT_DIFF_RADIANCE_HITDIST, od)
    den.set_user_texture(api.ResourceType.OUT_SPEC_RADIANCE_HITDIST, os_)
    return den, od, os_


def feed(den, frame):
    for k, v in frame.items():
        if not k.startswith("_"):
            den.set_user_texture(getattr(api.ResourceType, k), v)


def test_generator_and_oracle_reproduce_golden_fixture():
    g = torch.load(GOLDEN)
    w, h = g["width"], g["height"]
    for mode, robust in (("faithful", False), ("strict", True)):
        den, od, os_ = make(w, h, robust_mirror_test=robust)
        for f, stored in enumerate(g["inputs"]):
            fresh = synth.reblur_frame(f, w, h)
            for k in stored:
                assert torch.equal(fresh[k].view(torch.uint8), stored[k].view(torch.uint8)), f"generator drifted: frame {f} {k}"
            feed(den, {k: v.clone() for k, v in stored.items()})
            den.denoise(synth.common_settings(f, w, h))
            for got, want, name in ((od, g[mode][f][0], "diff"), (os_, g[mode][f][1], "spec")):
                r = compare(got, want, F16, atol=1e-4, rtol=2 ** -10)
                assert r["frac_bad"] == 0.0, f"{mode} frame {f} {name}: {r}"


def psnr(a, b, mask):
    mse = ((a - b) ** 2)[mask].mean().item()
    return 10 * torch.log10(torch.tensor(1.0 / max(mse, 1e-12))).item()


def test_denoising_improves_psnr_and_keeps_invariants():
    w, h, n = 160, 96, 10
    den, od, os_ = make(w, h)
    tm = lambda x: x / (1 + x)  # noqa: E731
    for f in range(n):
        fr = synth.reblur_frame(f, w, h, with_clean=True)
        feed(den, fr)
        den.denoise(synth.common_settings(f, w, h))
    m = fr["_hit"]
    for out, noisy_key, clean_key in ((od, "IN_DIFF_RADIANCE_HITDIST", "_clean_diff"), (os_, "IN_SPEC_RADIANCE_HITDIST", "_clean_spec")):
        clean = tm(fr[clean_key])
        p_noisy = psnr(tm(synth.unpack_radiance(fr[noisy_key])), clean, m)
        p_out = psnr(tm(synth.unpack_radiance(out)), clean, m)
        assert p_out > p_noisy + 8.0, (p_noisy, p_out)
        o = out.float()
        assert torch.isfinite(o).all()
        assert (o[..., 0] >= 0).all(), "luma (Y of YCoCg) must be non-negative"
        assert (o[..., 3] >= 0).all() and (o[..., 3] <= 1).all(), "normalized hit distance stays in [0, 1]"
    # history length is capped by maxAccumulatedFrameNum and grows by one per frame on static surfaces
    idata = den.textures[(int(api.ResourceType.PERMANENT_POOL), 2)].to(torch.int32) & 0xFFFF
    diff_frames = (idata & 63)[m].float()
    assert diff_frames.max() <= 30 and diff_frames.mean() > n - 3


def test_constant_input_gives_constant_output():
    """All weights normalise: a constant radiance field over a static camera stays constant."""
    w, h = 128, 80
    den, od, os_ = make(w, h)
    fr = synth.reblur_frame(0, w, h, with_clean=True)
    const = torch.zeros(h, w, 4, dtype=torch.float16)
    const[..., 0], const[..., 1], const[..., 2], const[..., 3] = 0.5, 0.125, -0.0625, 0.25
    const[~fr["_hit"]] = 0
    fr["IN_DIFF_RADIANCE_HITDIST"] = const.clone()
    fr["IN_SPEC_RADIANCE_HITDIST"] = const.clone()
    fr["IN_MV"] = torch.zeros_like(fr["IN_MV"])
    for f in range(6):
        feed(den, fr)
        cs = synth.common_settings(0, w, h)
        cs.frameIndex = f
        den.denoise(cs)
    m = fr["_hit"]
    for out in (od, os_):
        o = out.float()[m]
        assert (o[:, :3] - const.float()[m][:, :3]).abs().max() < 2e-3
        assert (o[:, 3] - 0.25).abs().max() < 2e-3


def test_restart_resets_history():
    w, h = 96, 64
    den, od, _ = make(w, h)
    for f in range(4):
        feed(den, synth.reblur_frame(f, w, h))
        den.denoise(synth.common_settings(f, w, h))
    fr = synth.reblur_frame(4, w, h, with_clean=True)
    feed(den, fr)
    den.denoise(synth.common_settings(4, w, h, accumulationMode=int(api.AccumulationMode.RESTART)))
    idata = den.textures[(int(api.ResourceType.PERMANENT_POOL), 2)].to(torch.int32) & 0xFFFF
    # gMaxAccumulatedFrameNum = 0 on reset frames (Reblur.cpp:364): min(n + 1, 0) -> history length 0 everywhere
    assert ((idata & 63)[fr["_hit"]] == 0).all()
    assert len(den.last_dispatches) == 7  # RESTART does not inject clears, CLEAR_AND_RESTART does


def test_hit_distance_reconstruction_fills_holes():
    """ReblurSettings::hitDistanceReconstructionMode (the NRD README's benchmark setting): with one lobe traced per pixel the other lobe's
    normalised hit distance arrives as 0; the 3x3 / 5x5 pass replaces it by a weighted average of same-surface neighbours and leaves the
    radiance untouched. Checked on the dispatch itself."""
    w, h = 128, 80
    for mode, shader_tail in ((1, "MODE_5X5=0"), (2, "MODE_5X5=1")):
        den, od, os_ = make(w, h)
        s = api.ReblurSettings(hitDistanceReconstructionMode=mode)
        seen = {}

        def before(i, d, keys, self):
            if d.shader.endswith(shader_tail) and "HitDistReconstruction" in d.shader:
                seen["in"] = [self.textures[k].clone() for k in keys]

        def after(i, d, keys, self):
            if d.shader.endswith(shader_tail) and "HitDistReconstruction" in d.shader:
                seen["out"] = [self.textures[k].clone() for k in keys]

        fr = synth.reblur_frame(0, w, h, holes=True)
        feed(den, fr)
        den.denoise(synth.common_settings(0, w, h), settings=s, before_dispatch=before, on_dispatch=after)
        assert "out" in seen, "the reconstruction pass must be in the dispatch list"
        viewz = seen["in"][2]
        inside = viewz < 5e5
        for lobe in (0, 1):
            src, dst = seen["in"][3 + lobe].float(), seen["out"][5 + lobe].float()
            holes = inside & (src[..., 3] == 0)
            assert holes.float().mean() > 0.2
            assert torch.equal(dst[..., :3][inside], src[..., :3][inside]), "radiance passes through"
            filled = (dst[..., 3][holes] > 0).float().mean().item()
            assert filled > 0.97, f"mode {mode} lobe {lobe}: only {filled:.3f} of the holes got a hit distance"
            # where data existed the result stays within the local range (the centre tap weighs 1000)
            had = inside & (src[..., 3] > 0)
            assert (dst[..., 3][had] - src[..., 3][had]).abs().max() < 0.05
