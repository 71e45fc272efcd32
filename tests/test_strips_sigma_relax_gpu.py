"""Row ranges ( multi-GPU strips ) for SIGMA and RELAX, as a single-GPU emulation like tests/test_parity_at_baseline_sizes_gpu.py does for REBLUR at 4K:
two full-size texture sets, every pass computes rows [0, y) on set A and [y, H) on set B through nrdcuDispatchRows, the rows a strip does not own are
poisoned after every pass and only the derived aprons cross the seam ( tiling.exchange_halos_local: the row arithmetic and apron table of the multi-GPU
path ). Every strip must equal the whole-frame launch of the same kernels bit for bit."""
import ctypes as C

import pytest
import torch

from nrd_sample_b200 import nrd_api as api, synth, tiling

pytestmark = pytest.mark.gpu
RT = api.ResourceType
F16 = api.Format.RGBA16_SFLOAT
INPUT_FORMATS = {"IN_VIEWZ": api.Format.R32_SFLOAT, "IN_NORMAL_ROUGHNESS": api.Format.R10_G10_B10_A2_UNORM, "IN_PENUMBRA": api.Format.R16_SFLOAT, "IN_TRANSLUCENCY": api.Format.RGBA8_UNORM}

CASES = {
    "sigma": (api.Denoiser.SIGMA_SHADOW, lambda f, w, h: synth.sigma_frame(f, w, h), [("OUT_SHADOW_TRANSLUCENCY", api.Format.R8_UNORM)],
              lambda: api.SigmaSettings(lightDirection=(C.c_float * 3)(0.0, 0.0, 1.0))),
    "sigma_translucency": (api.Denoiser.SIGMA_SHADOW_TRANSLUCENCY, lambda f, w, h: synth.sigma_frame(f, w, h, translucency=True), [("OUT_SHADOW_TRANSLUCENCY", api.Format.RGBA8_UNORM)],
                           lambda: api.SigmaSettings(lightDirection=(C.c_float * 3)(0.0, 0.0, 1.0))),
    "relax_sh": (api.Denoiser.RELAX_DIFFUSE_SPECULAR_SH, lambda f, w, h: synth.relax_frame(f, w, h), [(n, F16) for n in ("OUT_DIFF_SH0", "OUT_DIFF_SH1", "OUT_SPEC_SH0", "OUT_SPEC_SH1")], None),
    "relax_diffuse": (api.Denoiser.RELAX_DIFFUSE, lambda f, w, h: {k: v for k, v in synth.relax_frame(f, w, h, sh=False).items() if "SPEC" not in k},
                      [("OUT_DIFF_RADIANCE_HITDIST", F16)], lambda: api.RelaxSettings(enableAntiFirefly=True, hitDistanceReconstructionMode=1)),
}


@pytest.fixture(scope="module")
def ex():
    from nrd_sample_b200 import executor
    assert torch.cuda.is_available()
    return executor


@pytest.mark.parametrize("which", list(CASES))
def test_two_strips_equal_the_whole_frame(ex, which):
    den_id, frame_fn, outputs, settings = CASES[which]
    w, h, frames = 416, 304, 3          # 19 tile rows: strips of 160 and 144 rows, both taller than the 64-row apron
    dev = "cuda:0"
    host = ex.host_library()
    strips = tiling.strip_rows(h, 2)
    table = tiling.derive_halo_table(host, den_id, w, h, settings=settings() if settings else None)
    inst = api.NrdInstance(host, [(0, den_id)])
    assert inst.result == api.Result.SUCCESS
    if settings:
        assert inst.set_denoiser_settings(0, settings()) == api.Result.SUCCESS
    perm, tran = inst.pools()

    def texture_set():
        t = {}
        for kind, pool in ((RT.PERMANENT_POOL, perm), (RT.TRANSIENT_POOL, tran)):
            for i, (fmt, ds) in enumerate(pool):
                t[(int(kind), i)] = (ex.alloc_texture(fmt, (w + ds - 1) // ds, (h + ds - 1) // ds, dev), fmt)
        for name, fmt in outputs:
            t[(int(getattr(RT, name)), 0)] = (ex.alloc_texture(fmt, w, h, dev), fmt)
        return t

    sets = [texture_set(), texture_set(), texture_set()]   # strip A, strip B, the whole frame in one launch per pass
    pools = (int(RT.PERMANENT_POOL), int(RT.TRANSIENT_POOL))
    # the tile passes of SIGMA run over the whole frame on every strip ( kernels/sigma.cu ): nothing of theirs is poisoned or traded
    whole_frame_passes = ("Classify tiles", "Smooth tiles") if which.startswith("sigma") else ("Classify tiles",)
    for f in range(frames):
        for k, v in frame_fn(f, w, h).items():
            rt = getattr(RT, k)
            for s_ in sets:
                s_[(int(rt), 0)] = (v.to(dev), INPUT_FORMATS.get(k, F16))
        assert inst.set_common_settings(synth.common_settings(f, w, h)) == api.Result.SUCCESS
        r, dispatches = inst.get_compute_dispatches([0])
        assert r == api.Result.SUCCESS
        for d in dispatches:
            keys = [(b.type, b.index) if b.type in pools else (b.type, 0) for b in d.bindings]
            for si, rows in ((0, strips[0]), (1, strips[1]), (2, None)):
                ex.dispatch(d.shader, d.constants, [ex.texture_of(*sets[si][k]) for k in keys], flags=ex.FLAG_QUAD_INTRINSICS, rows=rows)
            if d.shader.startswith("Clear") or d.name.split(" - ")[-1] in whole_frame_passes:
                continue
            planes, halos = [[], []], []
            for j, (b, k) in enumerate(zip(d.bindings, keys)):
                if b.descriptor != int(api.DescriptorType.STORAGE_TEXTURE) or b.type == int(RT.IN_MV):
                    continue
                halos.append(tiling.halo_rows_for(table, d.name, j))
                for si in (0, 1):
                    t = sets[si][k][0]
                    p = t.view(torch.uint8).view(t.shape[0], -1)
                    y0, y1, _ = tiling._scaled(strips[si], tiling.HALO_ROWS, p.shape[0], h)
                    p[:y0] = 0xFF
                    p[y1:] = 0xFF
                    planes[si].append(p)
            tiling.exchange_halos_local(planes, strips, h, halos)
        torch.cuda.synchronize()
        for name, fmt in outputs:
            key = (int(getattr(RT, name)), 0)
            whole = sets[2][key][0]
            for si, (y0, y1) in enumerate(strips):
                got = sets[si][key][0]
                a, b = got[y0:y1].contiguous().view(torch.uint8), whole[y0:y1].contiguous().view(torch.uint8)
                assert torch.equal(a, b), f"{which} frame {f} strip {si} {name}: {(a != b).sum().item()} bytes differ from the whole-frame launch"


def test_motion_beyond_the_apron_is_reported(ex):
    """A strip holds 64 rows of its neighbours; history fetched at pixel + motion further away than 62 rows reads rows that never arrived. The executor scans the
    strip's motion vectors every frame ( kernels/peer_halo.cu ) and keeps the worst overshoot: nrdcuTileGetMotionBound."""
    w, h = 416, 304
    L = ex.load()
    den = ex.CudaDenoiser(api.Denoiser.SIGMA_SHADOW, w, h)
    out = ex.alloc_texture(api.Format.R8_UNORM, w, h, "cuda:0")
    den.set_user_texture(RT.OUT_SHADOW_TRANSLUCENCY, out, api.Format.R8_UNORM)
    frame = synth.sigma_frame(0, w, h, device="cuda:0")
    den.set_denoiser_settings(api.SigmaSettings(lightDirection=(C.c_float * 3)(0.0, 0.0, 1.0)))

    def run(mv):
        frame["IN_MV"] = mv
        for k, v in frame.items():
            den.set_user_texture(getattr(RT, k), v, INPUT_FORMATS.get(k, F16))
        den.set_common_settings(synth.common_settings(0, w, h))
        s = torch.cuda.current_stream().cuda_stream
        ex._check(L.nrdcuDenoiseRows(den.ctx, den._ids, 1, C.c_void_p(s), 0, 160, ex.DISPATCH_CALLBACK(), None), "nrdcuDenoiseRows")
        torch.cuda.synchronize()
        bound, worst = C.c_uint32(), C.c_uint32()
        ex._check(L.nrdcuTileGetMotionBound(den.ctx, C.byref(bound), C.byref(worst)), "nrdcuTileGetMotionBound")
        return bound.value, worst.value

    cs = synth.common_settings(0, w, h)
    scale_y = float(cs.motionVectorScale[1])          # uv per unit of IN_MV.y
    calm = torch.zeros_like(frame["IN_MV"])
    calm[..., 1] = 30.0 / (scale_y * h)                # 30 rows down everywhere: inside the apron
    assert run(calm) == (62, 0)
    wild = calm.clone()
    wild[150, 200, 1] = 100.0 / (scale_y * h)          # one pixel of the strip [0, 160) looks 100 rows down: row 250.5, the apron ends at 160 + 62
    bound, worst = run(wild)
    assert bound == 62 and worst == 29, (bound, worst)   # ceil( 250.5 - 222 )
    wild[150, 200, 1] = 400.0 / (scale_y * h)          # beyond the frame: not fetched, not counted ( the earlier overshoot stays on record )
    assert run(wild) == (62, 29)
    den.close()
