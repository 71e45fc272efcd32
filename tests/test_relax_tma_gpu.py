"""The tiled a-trous kernels of RELAX ( strides 2 and 4 ) take their four raw fp16 planes through TMA ( cp.async.bulk.tensor.2d, kernels/relax.cu ). Texels of a box
outside the texture arrive as zeros instead of clamped copies, which no tap may ever read: the frames must equal, bit for bit, what the LDG-staged instantiation
( NRD_B200_RELAX_TMA=0 ) produces — at a size whose right / bottom CTAs hang over the rect, for every RELAX signal / mode, and with a sub-rect ( dynamic resolution )."""
import os

import pytest
import torch

from nrd_sample_b200 import nrd_api as api, synth

pytestmark = pytest.mark.gpu
RT = api.ResourceType
F16 = api.Format.RGBA16_SFLOAT
INPUT_FORMATS = {"IN_VIEWZ": api.Format.R32_SFLOAT, "IN_NORMAL_ROUGHNESS": api.Format.R10_G10_B10_A2_UNORM}
GUIDES = ("IN_VIEWZ", "IN_NORMAL_ROUGHNESS", "IN_MV")
# denoiser, { its input resource: the tensor of synth.relax_frame bound to it }, outputs
CASES = {
    "diff_spec_sh": (api.Denoiser.RELAX_DIFFUSE_SPECULAR_SH, {k: k for k in ("IN_DIFF_SH0", "IN_DIFF_SH1", "IN_SPEC_SH0", "IN_SPEC_SH1")},
                     ("OUT_DIFF_SH0", "OUT_DIFF_SH1", "OUT_SPEC_SH0", "OUT_SPEC_SH1")),
    "diff_spec": (api.Denoiser.RELAX_DIFFUSE_SPECULAR, {"IN_DIFF_RADIANCE_HITDIST": "IN_DIFF_SH0", "IN_SPEC_RADIANCE_HITDIST": "IN_SPEC_SH0"},
                  ("OUT_DIFF_RADIANCE_HITDIST", "OUT_SPEC_RADIANCE_HITDIST")),
    "diff_sh": (api.Denoiser.RELAX_DIFFUSE_SH, {"IN_DIFF_SH0": "IN_DIFF_SH0", "IN_DIFF_SH1": "IN_DIFF_SH1"}, ("OUT_DIFF_SH0", "OUT_DIFF_SH1")),
    "spec": (api.Denoiser.RELAX_SPECULAR, {"IN_SPEC_RADIANCE_HITDIST": "IN_SPEC_SH0"}, ("OUT_SPEC_RADIANCE_HITDIST",)),
}


@pytest.fixture(scope="module")
def ex():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from nrd_sample_b200 import executor
    return executor


def run(ex, den_id, inputs, outputs, w, h, frames, tma, rect=None):
    os.environ["NRD_B200_RELAX_TMA"] = "1" if tma else "0"
    try:
        den = ex.CudaDenoiser(den_id, w, h)
        outs = {name: ex.alloc_texture(F16, w, h, "cuda:0") for name in outputs}
        for name in outputs:
            den.set_user_texture(getattr(RT, name), outs[name], F16)
        got = []
        for f in range(frames):
            frame = synth.relax_frame(f, w, h, device="cuda:0")
            for k in GUIDES:
                den.set_user_texture(getattr(RT, k), frame[k], INPUT_FORMATS.get(k, F16))
            for k, src in inputs.items():
                den.set_user_texture(getattr(RT, k), frame[src], F16)
            cs = synth.common_settings(f, w, h)
            if rect is not None:
                cs.rectSize[0], cs.rectSize[1] = rect
                cs.rectSizePrev[0], cs.rectSizePrev[1] = rect
            den.set_common_settings(cs)
            den.denoise()
            torch.cuda.synchronize()
            got.append({k: t.clone() for k, t in outs.items()})
        den.close()
        return got
    finally:
        os.environ.pop("NRD_B200_RELAX_TMA", None)


@pytest.mark.parametrize("which", list(CASES))
def test_tma_staging_equals_ldg_staging(ex, which):
    den_id, inputs, outputs = CASES[which]
    w, h = 328, 204   # 10.25 x 25.5 CTAs: the last CTA column / row reaches past the rect, their boxes past the texture
    a = run(ex, den_id, inputs, outputs, w, h, 6, tma=True)
    b = run(ex, den_id, inputs, outputs, w, h, 6, tma=False)
    for f, (x, y) in enumerate(zip(a, b)):
        for name in x:
            assert torch.equal(x[name], y[name]), f"{which} frame {f} {name}: TMA-staged a-trous differs from the LDG-staged one"
    assert float(a[-1][outputs[0]].float().abs().sum()) > 0.0


def test_tma_staging_with_a_sub_rect(ex):
    den_id, inputs, outputs = CASES["diff_spec_sh"]
    w, h = 320, 192
    a = run(ex, den_id, inputs, outputs, w, h, 5, tma=True, rect=(250, 150))
    b = run(ex, den_id, inputs, outputs, w, h, 5, tma=False, rect=(250, 150))
    for f, (x, y) in enumerate(zip(a, b)):
        for name in x:
            assert torch.equal(x[name], y[name]), f"frame {f} {name}: TMA-staged a-trous differs from the LDG-staged one with rectSize < resourceSize"
