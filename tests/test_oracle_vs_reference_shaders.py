"""Pins the CPU oracle's pixel math against the REFERENCE'S OWN SHADERS: External/NRD/Shaders/*.cs.hlsl (+ MathLib's ml.hlsli) compiled as
C++ from where they lie under /root/reference into oracle/_ref/libnrd_refshaders.so (oracle/ref_build_shaders.py, oracle/ref_shim/hlsl_cpu.h).
Every dispatch of every frame is replayed by both engines from the same pre-dispatch textures; every written texture must be bit-identical
(both are fp32 CPU code without FMA contraction, so there is no tolerance to hide behind).

The prebuilt .so travels to the GPU box; when it is absent and the reference tree is not mounted the tests skip and the committed fixtures
(tests/golden/refshaders_*.pt, written by tests/golden/make_refshader_golden.py from the same engine) take over."""
import ctypes as C
import os

import pytest
import torch

from nrd_sample_b200 import nrd_api as api, synth
from oracle import runner

RT = api.ResourceType
GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

def one_lobe(lobe):
    """synth.reblur_frame without the inputs of the other lobe ( REBLUR_DIFFUSE / REBLUR_SPECULAR )"""
    other = "SPEC" if lobe == "DIFF" else "DIFF"
    return lambda *a, **k: {key: v for key, v in synth.reblur_frame(*a, **k).items() if f"_{other}_" not in key}


DENOISERS = {
    "reblur": (api.Denoiser.REBLUR_DIFFUSE_SPECULAR, synth.reblur_frame, ("OUT_DIFF_RADIANCE_HITDIST", "OUT_SPEC_RADIANCE_HITDIST")),
    "reblur_diff": (api.Denoiser.REBLUR_DIFFUSE, one_lobe("DIFF"), ("OUT_DIFF_RADIANCE_HITDIST",)),
    "reblur_spec": (api.Denoiser.REBLUR_SPECULAR, one_lobe("SPEC"), ("OUT_SPEC_RADIANCE_HITDIST",)),
    "sigma": (api.Denoiser.SIGMA_SHADOW, synth.sigma_frame, ("OUT_SHADOW_TRANSLUCENCY",)),
    "relax": (api.Denoiser.RELAX_DIFFUSE_SPECULAR_SH, synth.relax_frame, ("OUT_DIFF_SH0", "OUT_DIFF_SH1", "OUT_SPEC_SH0", "OUT_SPEC_SH1")),
    "reference": (api.Denoiser.REFERENCE, synth.reference_frame, ("OUT_SIGNAL",)),
    "sigma_tr": (api.Denoiser.SIGMA_SHADOW_TRANSLUCENCY, lambda *a, **k: synth.sigma_frame(*a, translucency=True, **k), ("OUT_SHADOW_TRANSLUCENCY",)),
}
# user-texture formats that differ from runner.USER_FORMATS for a denoiser
OUTPUT_FORMATS = {"sigma_tr": {"OUT_SHADOW_TRANSLUCENCY": api.Format.RGBA8_UNORM}}
# denoisers the CUDA executor covers beyond the three the oracle restates: the reference shaders are their only CPU engine
REFERENCE_ONLY = {
    "relax_nosh": (api.Denoiser.RELAX_DIFFUSE_SPECULAR, lambda *a, **k: synth.relax_frame(*a, sh=False, **k), ("OUT_DIFF_RADIANCE_HITDIST", "OUT_SPEC_RADIANCE_HITDIST")),
}

# ( label, denoiser, width, height, frames, denoiser settings, extra frame kwargs )
CASES = [
    ("reblur_default", "reblur", 96, 64, 6, None, {}),
    ("reblur_odd_size", "reblur", 100, 75, 4, None, {}),
    ("reblur_recon3x3", "reblur", 96, 64, 4, lambda: api.ReblurSettings(hitDistanceReconstructionMode=1), {"holes": True}),
    ("reblur_recon5x5", "reblur", 96, 64, 3, lambda: api.ReblurSettings(hitDistanceReconstructionMode=2), {"holes": True}),
    ("reblur_no_stabilization", "reblur", 96, 64, 4, lambda: api.ReblurSettings(maxStabilizedFrameNum=0), {}),
    ("reblur_no_prepass_no_antifirefly", "reblur", 96, 64, 4, lambda: api.ReblurSettings(diffusePrepassBlurRadius=0.0, specularPrepassBlurRadius=0.0, enableAntiFirefly=False), {}),
    ("reblur_checkerboard_white", "reblur", 96, 64, 5, lambda: api.ReblurSettings(checkerboardMode=2), {"checkerboard": 2}),
    ("reblur_checkerboard_black_no_prepass_radius", "reblur", 100, 76, 4, lambda: api.ReblurSettings(checkerboardMode=1, diffusePrepassBlurRadius=0.0, specularPrepassBlurRadius=0.0),
     {"checkerboard": 1}),
    # "cs_*" keys go to CommonSettings, the rest to the frame generator
    ("reblur_confidence_and_threshold_mix", "reblur", 96, 64, 5, None, {"guides": True, "cs_isHistoryConfidenceAvailable": True, "cs_isDisocclusionThresholdMixAvailable": True}),
    ("reblur_confidence_checkerboard", "reblur", 96, 64, 4, lambda: api.ReblurSettings(checkerboardMode=2), {"guides": True, "checkerboard": 2, "cs_isHistoryConfidenceAvailable": True}),
    ("reblur_split_screen", "reblur", 96, 64, 3, None, {"cs_splitScreen": 0.4}),
    ("reblur_split_screen_checkerboard_full", "reblur", 96, 64, 2, lambda: api.ReblurSettings(checkerboardMode=1), {"checkerboard": 1, "cs_splitScreen": 1.0}),
    ("reblur_diffuse_default", "reblur_diff", 96, 64, 5, None, {}),
    ("reblur_diffuse_recon3x3_no_stabilization_odd_size", "reblur_diff", 100, 75, 4, lambda: api.ReblurSettings(hitDistanceReconstructionMode=1, maxStabilizedFrameNum=0), {"holes": True}),
    ("reblur_diffuse_checkerboard_guides_split", "reblur_diff", 96, 64, 4, lambda: api.ReblurSettings(checkerboardMode=2),
     {"guides": True, "checkerboard": 2, "cs_isHistoryConfidenceAvailable": True, "cs_isDisocclusionThresholdMixAvailable": True, "cs_splitScreen": 0.4}),
    ("reblur_specular_default", "reblur_spec", 96, 64, 5, None, {}),
    ("reblur_specular_recon5x5_no_stabilization_odd_size", "reblur_spec", 100, 75, 4, lambda: api.ReblurSettings(hitDistanceReconstructionMode=2, maxStabilizedFrameNum=0), {"holes": True}),
    ("reblur_specular_checkerboard_guides_no_prepass", "reblur_spec", 96, 64, 4, lambda: api.ReblurSettings(checkerboardMode=1, specularPrepassBlurRadius=0.0),
     {"guides": True, "checkerboard": 1, "cs_isHistoryConfidenceAvailable": True, "cs_isDisocclusionThresholdMixAvailable": True}),
    # "cs_static": the camera of frame 0 every frame, so the REFERENCE accumulator actually accumulates ( Reference.hpp:62-68 )
    ("reference_static_camera_split", "reference", 100, 75, 5, lambda: api.ReferenceSettings(maxAccumulatedFrameNum=3), {"cs_static": True, "cs_splitScreen": 0.3}),
    ("reference_moving_camera", "reference", 96, 64, 3, None, {}),
    ("sigma_default", "sigma", 96, 64, 5, None, {}),
    ("sigma_odd_size", "sigma", 100, 75, 3, None, {}),
    ("sigma_no_stabilization", "sigma", 96, 64, 3, lambda: api.SigmaSettings(lightDirection=(C.c_float * 3)(0.3, 0.8, -0.5), maxStabilizedFrameNum=0), {}),
    ("sigma_translucency_default", "sigma_tr", 96, 64, 5, None, {}),
    ("sigma_translucency_odd_size_no_stabilization", "sigma_tr", 100, 75, 3, lambda: api.SigmaSettings(lightDirection=(C.c_float * 3)(0.3, 0.8, -0.5), maxStabilizedFrameNum=0), {}),
    ("relax_default", "relax", 96, 64, 5, None, {}),
    ("relax_recon3x3", "relax", 96, 64, 3, lambda: api.RelaxSettings(hitDistanceReconstructionMode=1), {"holes": True}),
    ("relax_recon5x5_odd_size_no_prepass", "relax", 100, 75, 3, lambda: api.RelaxSettings(hitDistanceReconstructionMode=2, diffusePrepassBlurRadius=0.0, specularPrepassBlurRadius=0.0),
     {"holes": True}),
    ("relax_split_screen_checkerboard", "relax", 96, 64, 3, lambda: api.RelaxSettings(checkerboardMode=2), {"checkerboard": 2, "cs_splitScreen": 0.45}),
    ("relax_checkerboard_white", "relax", 96, 64, 4, lambda: api.RelaxSettings(checkerboardMode=2), {"checkerboard": 2}),
    ("relax_checkerboard_black_guides", "relax", 100, 76, 4, lambda: api.RelaxSettings(checkerboardMode=1, enableAntiFirefly=True),
     {"checkerboard": 1, "guides": True, "cs_isHistoryConfidenceAvailable": True, "cs_isDisocclusionThresholdMixAvailable": True}),
    ("relax_odd_size", "relax", 100, 75, 3, None, {}),
    ("relax_antifirefly_3_iterations", "relax", 96, 64, 4, lambda: api.RelaxSettings(enableAntiFirefly=True, atrousIterationNum=3), {}),
    ("relax_8_iterations_no_prepass", "relax", 96, 64, 3, lambda: api.RelaxSettings(atrousIterationNum=8, diffusePrepassBlurRadius=0.0, specularPrepassBlurRadius=0.0,
                                                                                  enableRoughnessEdgeStopping=False), {}),
]


def make_denoiser(which, w, h, engine="oracle"):
    den_id, _, outputs = {**DENOISERS, **REFERENCE_ONLY}[which]
    den = runner.OracleDenoiser(runner.default_host_library(), den_id, w, h, engine=engine)
    for o in outputs:
        fmt = OUTPUT_FORMATS.get(which, {}).get(o, runner.USER_FORMATS[getattr(RT, o)])
        den.set_user_texture(getattr(RT, o), runner.alloc_texture(fmt, w, h), fmt)
    return den


needs_refshaders = pytest.mark.skipif(runner.ref_shaders() is None, reason="oracle/_ref/libnrd_refshaders.so not built (reference tree not mounted)")


@needs_refshaders
def test_every_shader_the_host_library_emits_is_compiled_from_the_reference():
    names = set(runner.ref_shader_names())
    for which, (den_id, _, _) in DENOISERS.items():
        inst = api.NrdInstance(runner.default_host_library(), [(0, den_id)])
        for s in set(inst.shader_identifiers()):
            if "Validation" in s:
                continue   # debug overlay, out of scope
            assert s in names, f"{s} is not in libnrd_refshaders.so"


@needs_refshaders
@pytest.mark.parametrize("label,which,w,h,frames,make_settings,frame_kwargs", CASES, ids=[c[0] for c in CASES])
def test_oracle_is_bit_identical_to_the_reference_shaders_per_dispatch(label, which, w, h, frames, make_settings, frame_kwargs):
    den = make_denoiser(which, w, h)
    fn = runner.ref_shaders().nrd_refshader_dispatch
    snap, checked, seen = {}, [0], set()

    def before(i, d, keys, self):
        snap["t"] = [self.textures[k].clone() for k in keys]

    def after(i, d, keys, self):
        ref = snap["t"]
        arr = (runner.OracleTexture * len(keys))(*[runner.tex_desc(t, self.formats[k]) for t, k in zip(ref, keys)])
        cb = C.create_string_buffer(d.constants, len(d.constants)) if d.constants else None
        rc = fn(d.shader.encode(), cb, len(d.constants), arr, len(keys), d.grid[0], d.grid[1], 0)
        assert rc == 0, f"reference shader {d.shader} rc={rc}"
        seen.add(d.shader)
        for j, (b, k) in enumerate(zip(d.bindings, keys)):
            if b.descriptor != int(api.DescriptorType.STORAGE_TEXTURE):
                assert torch.equal(ref[j], self.textures[k]), f"{d.name}: input {j} was modified"
                continue
            same = torch.equal(ref[j], self.textures[k])
            if not same:
                a, o = ref[j].float().flatten(), self.textures[k].float().flatten()
                bad = (ref[j] != self.textures[k]).float().mean().item()
                raise AssertionError(f"{label} frame {frame} {d.name} ({d.shader}) output {j}: {bad:.2e} of the texels differ, max |d| {(a - o).abs().max().item():.3g}")
            checked[0] += 1

    settings = make_settings() if make_settings else None
    cs_kwargs = {k[3:]: v for k, v in frame_kwargs.items() if k.startswith("cs_")}
    frame_kwargs = {k: v for k, v in frame_kwargs.items() if not k.startswith("cs_")}
    for frame in range(frames):
        for k, v in DENOISERS[which][1](frame, w, h, **frame_kwargs).items():
            den.set_user_texture(getattr(RT, k), v)
        static = cs_kwargs.get("static", False)
        cs = synth.common_settings(0 if static else frame, w, h, **{k: v for k, v in cs_kwargs.items() if k != "static"})
        cs.frameIndex = frame
        den.denoise(cs, settings=settings, before_dispatch=before, on_dispatch=after)
    few = which == "reference" or cs_kwargs.get("splitScreen", 0.0) >= 1.0    # chains of one or two passes
    assert checked[0] >= frames * (1 if few else 5)
    assert any("Clear" in s for s in seen) and len(seen) >= (2 if few else 6)


@needs_refshaders
@pytest.mark.parametrize("which", ["reblur", "reblur_diff", "reblur_spec", "sigma", "relax", "sigma_tr"])
def test_closed_loop_with_the_reference_shaders_as_the_engine(which):
    """The whole recurrence (history feedback) executed by the reference's shaders, against the oracle: final outputs identical."""
    w, h, frames = 112, 80, 5
    a, b = make_denoiser(which, w, h, "oracle"), make_denoiser(which, w, h, "reference")
    for frame in range(frames):
        fr = DENOISERS[which][1](frame, w, h)
        cs = synth.common_settings(frame, w, h)
        for den in (a, b):
            for k, v in fr.items():
                den.set_user_texture(getattr(RT, k), v)
            den.denoise(cs)
        for o in DENOISERS[which][2]:
            key = (int(getattr(RT, o)), 0)
            assert torch.equal(a.textures[key], b.textures[key]), f"{which} frame {frame} {o}"


@pytest.mark.parametrize("which", ["reblur", "sigma", "relax"])
def test_oracle_reproduces_the_committed_reference_shader_outputs(which):
    """Runs everywhere (GPU box included): the fixture holds the outputs the reference's shaders produced in this container."""
    path = os.path.join(GOLDEN_DIR, f"refshaders_{which}_80x48.pt")
    if not os.path.exists(path):
        pytest.skip("fixture not generated")
    g = torch.load(path)
    w, h = g["width"], g["height"]
    den = make_denoiser(which, w, h)
    for frame in range(g["frames"]):
        for k, v in DENOISERS[which][1](frame, w, h).items():
            den.set_user_texture(getattr(RT, k), v)
        den.denoise(synth.common_settings(frame, w, h))
    for o, ref in g["outputs"].items():
        assert torch.equal(den.textures[(int(getattr(RT, o)), 0)], ref), f"{which} {o} differs from the reference-shader fixture"
