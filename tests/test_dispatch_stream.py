"""Pins the host pass graph against the reference host library (External/NRD/Source/*.cpp compiled unmodified into
oracle/_ref/libnrd_ref.so): InstanceDesc, pools, shader identifiers, per-frame dispatch order, bindings, ping-pong parity,
grids, constant-buffer bytes. When /root/reference is not mounted (GPU box) the same streams are checked against the
JSON snapshot tests/golden/dispatch_streams.json, which tests/golden/make_dispatch_golden.py wrote from the reference build."""
import ctypes as C
import json
import os
import struct

import numpy as np
import pytest

from nrd_sample_b200 import nrd_api as api, synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dispatch_streams.json")

CASES = [
    ("reblur_1080p", api.Denoiser.REBLUR_DIFFUSE_SPECULAR, 1920, 1080, None),
    ("reblur_1440p", api.Denoiser.REBLUR_DIFFUSE_SPECULAR, 2560, 1440, None),
    ("reblur_720p_recon_nots", api.Denoiser.REBLUR_DIFFUSE_SPECULAR, 1280, 720, "recon_nots"),
    ("reblur_odd_noprepass", api.Denoiser.REBLUR_DIFFUSE_SPECULAR, 1000, 562, "noprepass"),
    ("reblur_checkerboard_guides_split", api.Denoiser.REBLUR_DIFFUSE_SPECULAR, 1280, 720, "reblur_cb_guides_split"),
    ("reblur_split_only", api.Denoiser.REBLUR_DIFFUSE_SPECULAR, 640, 360, "split_only"),
    ("reference_720p", api.Denoiser.REFERENCE, 1280, 720, "reference"),
    ("reference_static_camera", api.Denoiser.REFERENCE, 640, 360, "reference_static"),
    ("reblur_diffuse_1080p", api.Denoiser.REBLUR_DIFFUSE, 1920, 1080, None),
    ("reblur_diffuse_recon_nots", api.Denoiser.REBLUR_DIFFUSE, 1000, 562, "recon_nots"),
    ("reblur_specular_1080p", api.Denoiser.REBLUR_SPECULAR, 1920, 1080, None),
    ("reblur_specular_cb_guides_split", api.Denoiser.REBLUR_SPECULAR, 1280, 720, "reblur_cb_guides_split"),
    ("reblur_specular_noprepass", api.Denoiser.REBLUR_SPECULAR, 640, 360, "noprepass"),
    ("reblur_sh_1440p", api.Denoiser.REBLUR_DIFFUSE_SPECULAR_SH, 2560, 1440, None),
    ("reblur_sh_cb_guides_split", api.Denoiser.REBLUR_DIFFUSE_SPECULAR_SH, 1280, 720, "reblur_cb_guides_split"),
    ("reblur_sh_recon_nots", api.Denoiser.REBLUR_DIFFUSE_SPECULAR_SH, 1000, 562, "recon_nots"),
    ("reblur_diffuse_sh_1080p", api.Denoiser.REBLUR_DIFFUSE_SH, 1920, 1080, None),
    ("reblur_diffuse_sh_noprepass", api.Denoiser.REBLUR_DIFFUSE_SH, 640, 360, "noprepass"),
    ("reblur_specular_sh_recon_nots", api.Denoiser.REBLUR_SPECULAR_SH, 1280, 720, "recon_nots"),
    ("reblur_specular_sh_cb_guides_split", api.Denoiser.REBLUR_SPECULAR_SH, 1000, 562, "reblur_cb_guides_split"),
    # NRD_MODE = OCCLUSION ( hit distance only: R16_UNORM / R8_UNORM pools, no pre-pass, no stabilization ) and NRD_MODE = DO ( RGBA16_SNORM )
    ("reblur_occlusion_1080p", api.Denoiser.REBLUR_DIFFUSE_SPECULAR_OCCLUSION, 1920, 1080, None),
    ("reblur_occlusion_cb_guides_split", api.Denoiser.REBLUR_DIFFUSE_SPECULAR_OCCLUSION, 1280, 720, "reblur_cb_guides_split"),
    ("reblur_occlusion_recon", api.Denoiser.REBLUR_DIFFUSE_SPECULAR_OCCLUSION, 1000, 562, "recon_nots"),
    ("reblur_diffuse_occlusion_720p", api.Denoiser.REBLUR_DIFFUSE_OCCLUSION, 1280, 720, None),
    ("reblur_diffuse_occlusion_recon", api.Denoiser.REBLUR_DIFFUSE_OCCLUSION, 640, 360, "recon_nots"),
    ("reblur_specular_occlusion_720p", api.Denoiser.REBLUR_SPECULAR_OCCLUSION, 1280, 720, None),
    ("reblur_specular_occlusion_cb_guides_split", api.Denoiser.REBLUR_SPECULAR_OCCLUSION, 1000, 562, "reblur_cb_guides_split"),
    ("reblur_directional_occlusion_1080p", api.Denoiser.REBLUR_DIFFUSE_DIRECTIONAL_OCCLUSION, 1920, 1080, None),
    ("reblur_directional_occlusion_recon_nots", api.Denoiser.REBLUR_DIFFUSE_DIRECTIONAL_OCCLUSION, 1000, 562, "recon_nots"),
    ("reblur_directional_occlusion_cb_guides_split", api.Denoiser.REBLUR_DIFFUSE_DIRECTIONAL_OCCLUSION, 1280, 720, "reblur_cb_guides_split"),
    # CommonSettings::enableValidation: one more dispatch at the end of the frame ( REBLUR_ADD_VALIDATION_DISPATCH / RELAX_ADD_VALIDATION_DISPATCH )
    ("reblur_validation", api.Denoiser.REBLUR_DIFFUSE_SPECULAR, 1280, 720, "validation"),
    ("reblur_specular_validation", api.Denoiser.REBLUR_SPECULAR, 640, 360, "validation"),
    ("reblur_diffuse_sh_validation", api.Denoiser.REBLUR_DIFFUSE_SH, 640, 360, "validation"),
    ("reblur_occlusion_validation", api.Denoiser.REBLUR_DIFFUSE_SPECULAR_OCCLUSION, 1000, 562, "validation"),
    ("reblur_directional_occlusion_validation", api.Denoiser.REBLUR_DIFFUSE_DIRECTIONAL_OCCLUSION, 640, 360, "validation"),
    ("relax_validation", api.Denoiser.RELAX_DIFFUSE_SPECULAR, 1280, 720, "validation"),
    ("relax_specular_sh_validation", api.Denoiser.RELAX_SPECULAR_SH, 640, 360, "validation"),
    ("sigma_512", api.Denoiser.SIGMA_SHADOW, 512, 512, "sigma"),
    ("sigma_nostab", api.Denoiser.SIGMA_SHADOW, 640, 360, "sigma_nostab"),
    ("sigma_translucency_1080p", api.Denoiser.SIGMA_SHADOW_TRANSLUCENCY, 1920, 1080, "sigma"),
    ("sigma_translucency_nostab", api.Denoiser.SIGMA_SHADOW_TRANSLUCENCY, 1000, 562, "sigma_nostab"),
    ("relax_sh_1440p", api.Denoiser.RELAX_DIFFUSE_SPECULAR_SH, 2560, 1440, None),
    ("relax_sh_odd_firefly_recon", api.Denoiser.RELAX_DIFFUSE_SPECULAR_SH, 1000, 562, "relax_firefly_recon"),
    ("relax_sh_8_iterations", api.Denoiser.RELAX_DIFFUSE_SPECULAR_SH, 640, 360, "relax_8"),
    ("relax_1440p", api.Denoiser.RELAX_DIFFUSE_SPECULAR, 2560, 1440, None),
    ("relax_odd_firefly_recon", api.Denoiser.RELAX_DIFFUSE_SPECULAR, 1000, 562, "relax_firefly_recon"),
    ("relax_diffuse_1080p", api.Denoiser.RELAX_DIFFUSE, 1920, 1080, None),
    ("relax_diffuse_firefly_recon", api.Denoiser.RELAX_DIFFUSE, 1000, 562, "relax_firefly_recon"),
    ("relax_diffuse_sh_cb_guides_split", api.Denoiser.RELAX_DIFFUSE_SH, 1280, 720, "relax_cb_guides_split"),
    ("relax_diffuse_sh_8_iterations", api.Denoiser.RELAX_DIFFUSE_SH, 640, 360, "relax_8"),
    ("relax_specular_1080p", api.Denoiser.RELAX_SPECULAR, 1920, 1080, None),
    ("relax_specular_cb_guides_split", api.Denoiser.RELAX_SPECULAR, 1280, 720, "relax_cb_guides_split"),
    ("relax_specular_sh_firefly_recon", api.Denoiser.RELAX_SPECULAR_SH, 1000, 562, "relax_firefly_recon"),
    ("relax_specular_sh_8_iterations", api.Denoiser.RELAX_SPECULAR_SH, 640, 360, "relax_8"),
]


# CommonSettings overrides per case kind; "static" freezes the camera so the REFERENCE frame counter advances
COMMON = {
    "reblur_cb_guides_split": dict(isHistoryConfidenceAvailable=True, isDisocclusionThresholdMixAvailable=True, splitScreen=0.4),
    "relax_cb_guides_split": dict(isHistoryConfidenceAvailable=True, isDisocclusionThresholdMixAvailable=True, splitScreen=0.4),
    "split_only": dict(splitScreen=1.0),
    "validation": dict(enableValidation=True),
    "reference": dict(splitScreen=0.25),
}


def _settings(kind):
    if kind == "reblur_cb_guides_split":
        return api.ReblurSettings(checkerboardMode=2)
    if kind == "reference_static":
        return api.ReferenceSettings(maxAccumulatedFrameNum=2)
    if kind == "recon_nots":
        return api.ReblurSettings(hitDistanceReconstructionMode=1, maxStabilizedFrameNum=0)
    if kind == "noprepass":
        return api.ReblurSettings(diffusePrepassBlurRadius=0.0, specularPrepassBlurRadius=0.0, enableAntiFirefly=False)
    if kind == "sigma":
        return api.SigmaSettings(lightDirection=(C.c_float * 3)(0.0, 0.0, 1.0))
    if kind == "sigma_nostab":
        return api.SigmaSettings(lightDirection=(C.c_float * 3)(0.3, 0.8, -0.5), maxStabilizedFrameNum=0)
    if kind == "relax_firefly_recon":
        return api.RelaxSettings(enableAntiFirefly=True, hitDistanceReconstructionMode=2, atrousIterationNum=4, diffuseMinLuminanceWeight=0.1, specularLobeAngleSlack=0.3,
                                 diffuseMaxAccumulatedFrameNum=40, historyFixFrameNum=2)
    if kind == "relax_cb_guides_split":
        return api.RelaxSettings(checkerboardMode=1, enableAntiFirefly=True)
    if kind == "relax_8":
        return api.RelaxSettings(atrousIterationNum=8, diffusePrepassBlurRadius=0.0, specularPrepassBlurRadius=0.0, enableRoughnessEdgeStopping=False)
    return None


def capture(lib, denoiser, w, h, kind, frames=4):
    """Everything observable through the descriptor API, as plain python."""
    inst = api.NrdInstance(lib, [(7, denoiser)])
    assert inst.result == api.Result.SUCCESS
    d = inst.desc()
    pd = d.descriptorPoolDesc
    out = {
        "desc": [d.constantBufferAndSamplersSpaceIndex, d.resourcesSpaceIndex, d.constantBufferRegisterIndex, d.samplersBaseRegisterIndex, d.resourcesBaseRegisterIndex,
                 d.constantBufferMaxDataSize, d.samplersNum, d.pipelinesNum, d.permanentPoolSize, d.transientPoolSize],
        "descriptor_pool": [pd.perSetTexturesMaxNum, pd.perSetStorageTexturesMaxNum, pd.totalTexturesNum, pd.totalStorageTexturesNum, pd.setsMaxNum],
        "pools": [list(map(list, p)) for p in inst.pools()],
        "shaders": inst.shader_identifiers(),
        "ranges": [[(d.pipelines[i].resourceRanges[j].descriptorType, d.pipelines[i].resourceRanges[j].descriptorsNum) for j in range(d.pipelines[i].resourceRangesNum)]
                   for i in range(d.pipelinesNum)],
        "frames": [],
    }
    out["ranges"] = [[list(x) for x in r] for r in out["ranges"]]
    s = _settings(kind)
    for f in range(frames):
        cs = synth.common_settings(0 if kind == "reference_static" else f, w, h, **COMMON.get(kind, {}))
        cs.frameIndex = f
        if f == 2 and kind != "reference_static":
            cs.cameraJitter = (C.c_float * 2)(0.25, -0.125)
        r1 = inst.set_common_settings(cs)
        if s is not None:
            inst.set_denoiser_settings(7, s)
        r2, disp = inst.get_compute_dispatches([7])
        out["frames"].append({"results": [int(r1), int(r2)], "dispatches": [
            {"name": x.name, "shader": x.shader, "bindings": [[b.descriptor, b.type, b.index] for b in x.bindings], "grid": list(x.grid),
             "pipeline": x.pipeline_index, "cb_same_as_prev": x.constants_match_previous, "cb": x.constants.hex()} for x in disp]})
    return out


def assert_same(mine, ref, label):
    for key in ("desc", "descriptor_pool", "pools", "shaders", "ranges"):
        assert mine[key] == ref[key], f"{label}: {key} differs"
    assert len(mine["frames"]) == len(ref["frames"])
    for f, (fm, fr) in enumerate(zip(mine["frames"], ref["frames"])):
        assert fm["results"] == fr["results"], f"{label} frame {f}: result codes"
        assert len(fm["dispatches"]) == len(fr["dispatches"]), f"{label} frame {f}: dispatch count"
        for i, (a, b) in enumerate(zip(fm["dispatches"], fr["dispatches"])):
            for key in ("name", "shader", "bindings", "grid", "pipeline", "cb_same_as_prev"):
                assert a[key] == b[key], f"{label} frame {f} dispatch {i} ({b['name']}): {key}: {a[key]} != {b[key]}"
            ca, cbb = bytes.fromhex(a["cb"]), bytes.fromhex(b["cb"])
            assert len(ca) == len(cbb)
            if not ca:
                continue
            ia, ib = np.frombuffer(ca, np.uint32), np.frombuffer(cbb, np.uint32)
            fa, fb = np.frombuffer(ca, np.float32), np.frombuffer(cbb, np.float32)
            differ = np.nonzero(ia != ib)[0]
            # every constant-buffer word is bit-identical: integers, matrices, the rotators ( both sides evaluate cosf / sinf of the same fp32 angle ) and
            # even the never-read -0.0f in gViewVectorWorld.w that the reference's SSE negation leaves behind
            assert len(differ) == 0, f"{label} frame {f} dispatch {i} ({b['name']}): words {list(differ)} differ: {[(float(fa[k]), float(fb[k])) for k in differ]}"


@pytest.mark.parametrize("label,denoiser,w,h,kind", CASES)
def test_stream_matches_reference_library(host_library, reference_host_library, label, denoiser, w, h, kind):
    assert_same(capture(host_library, denoiser, w, h, kind), capture(reference_host_library, denoiser, w, h, kind), label)


@pytest.mark.parametrize("label,denoiser,w,h,kind", CASES)
def test_stream_matches_golden_snapshot(host_library, label, denoiser, w, h, kind):
    if not os.path.exists(GOLDEN):
        pytest.skip("golden snapshot not generated")
    golden = json.load(open(GOLDEN))
    assert_same(capture(host_library, denoiser, w, h, kind), golden[label], label)
