"""CPU-side checks of the RELAX_DIFFUSE_SPECULAR_SH oracle (oracle/relax_passes.cpp) and of the synthetic SH input generator:
regression fixture, dispatch list, denoising quality and invariants (the pin against the reference's shaders is tests/test_oracle_vs_reference_shaders.py)."""
import os

import torch

from nrd_sample_b200 import nrd_api as api, synth
from oracle import runner
from tests.util import compare

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "relax_96x64.pt")
F16 = api.Format.RGBA16_SFLOAT
RT = api.ResourceType
OUTS = ("OUT_DIFF_SH0", "OUT_DIFF_SH1", "OUT_SPEC_SH0", "OUT_SPEC_SH1")


def make(w, h):
    den = runner.OracleDenoiser(runner.default_host_library(), api.Denoiser.RELAX_DIFFUSE_SPECULAR_SH, w, h)
    outs = {k: runner.alloc_texture(F16, w, h) for k in OUTS}
    for k, v in outs.items():
        den.set_user_texture(getattr(RT, k), v)
    return den, outs


def feed(den, frame):
    for k, v in frame.items():
        if not k.startswith("_"):
            den.set_user_texture(getattr(RT, k), v)


def ycocg_to_rgb(c):
    t = c[..., 0] - c[..., 2]
    return torch.stack([t + c[..., 1], c[..., 0] + c[..., 2], t - c[..., 1]], -1).clamp_min(0)


def test_generator_and_oracle_reproduce_golden_fixture():
    g = torch.load(GOLDEN)
    w, h = g["width"], g["height"]
    den, outs = make(w, h)
    for f, stored in enumerate(g["inputs"]):
        fresh = synth.relax_frame(f, w, h)
        for k in stored:
            assert torch.equal(fresh[k].view(torch.uint8), stored[k].view(torch.uint8)), f"generator drifted: frame {f} {k}"
        feed(den, {k: v.clone() for k, v in stored.items()})
        den.denoise(synth.common_settings(f, w, h))
        for k in OUTS:
            r = compare(outs[k], g["outputs"][f][k], F16, atol=1e-4, rtol=2 ** -10)
            assert r["frac_bad"] == 0.0, f"frame {f} {k}: {r}"


def test_dispatch_list():
    den, _ = make(96, 64)
    feed(den, synth.relax_frame(0, 96, 64))
    names = [d.shader.split("|")[0] for d in den.denoise(synth.common_settings(0, 96, 64))]
    chain = ["RELAX_ClassifyTiles.cs.hlsl", "RELAX_PrePass.cs.hlsl", "RELAX_TemporalAccumulation.cs.hlsl", "RELAX_HistoryFix.cs.hlsl", "RELAX_HistoryClamping.cs.hlsl",
             "RELAX_AtrousSmem.cs.hlsl"] + ["RELAX_Atrous.cs.hlsl"] * 4
    assert names[-10:] == chain and all(n.startswith("Clear") for n in names[:-10])
    # anti-firefly adds Copy + AntiFirefly, the iteration count is clamped to [2, 8]
    den.instance.set_denoiser_settings(den.identifier, api.RelaxSettings(enableAntiFirefly=True, atrousIterationNum=9))
    feed(den, synth.relax_frame(1, 96, 64))
    names = [d.shader.split("|")[0] for d in den.denoise(synth.common_settings(1, 96, 64))]
    assert names == chain[:5] + ["RELAX_Copy.cs.hlsl", "RELAX_AntiFirefly.cs.hlsl", "RELAX_AtrousSmem.cs.hlsl"] + ["RELAX_Atrous.cs.hlsl"] * 7


def test_denoising_improves_and_keeps_invariants():
    w, h, n = 160, 96, 10
    den, outs = make(w, h)
    tm = lambda x: x / (1 + x)  # noqa: E731
    for f in range(n):
        fr = synth.relax_frame(f, w, h, with_clean=(f == n - 1))
        feed(den, fr)
        den.denoise(synth.common_settings(f, w, h))
    m = fr["_hit"]
    for o, nk, ck in (("OUT_DIFF_SH0", "IN_DIFF_SH0", "_clean_diff"), ("OUT_SPEC_SH0", "IN_SPEC_SH0", "_clean_spec")):
        out = outs[o].float()
        assert torch.isfinite(out).all()
        clean = tm(fr[ck])
        mse_out = ((tm(ycocg_to_rgb(out[..., :3])) - clean) ** 2)[m].mean().item()
        mse_in = ((tm(fr[nk].float()[..., :3]) - clean) ** 2)[m].mean().item()
        assert mse_out * 10 < mse_in, (o, mse_in, mse_out)
        assert (out[..., 0][m] >= 0).all(), "Y of YCoCg is non-negative"
        hist = out[..., 3][m]   # the last a-trous pass stores history length - 1 in .w
        assert hist.max() <= n and hist.mean() > n - 3
    for o in ("OUT_DIFF_SH1", "OUT_SPEC_SH1"):
        assert torch.isfinite(outs[o].float()).all()
    # history length plane: 8-bit, frames / 255
    hl = den.textures[(int(RT.PERMANENT_POOL), 10)]
    assert hl.dtype == torch.uint8 and int(hl[m].max()) <= n + 1


def test_anti_firefly_and_eight_iterations_run():
    w, h = 96, 64
    den, outs = make(w, h)
    den.instance.set_denoiser_settings(den.identifier, api.RelaxSettings(enableAntiFirefly=True, atrousIterationNum=8))
    for f in range(3):
        feed(den, synth.relax_frame(f, w, h))
        den.denoise(synth.common_settings(f, w, h))
    for o in OUTS:
        assert torch.isfinite(outs[o].float()).all()
