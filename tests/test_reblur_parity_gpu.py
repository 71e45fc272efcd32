"""GPU parity of the CUDA REBLUR_DIFFUSE_SPECULAR path against the CPU oracle, through the C ABI (nrdcuDispatch /
nrdcuDenoise / nrdcuDenoiseHost). Tolerances (ours — the reference states none):
  strict mode (robust mirror predicate, DESIGN.md "chaotic predicates"): per pass, oracle-fed inputs:
      fp16/fp32 planes |a-b| <= 1e-3 + 2^-9 |b| on >= 99.9 % of texels and PSNR >= 60 dB; UNORM planes +-1 LSB; packed uints identical
      on >= 99 % of texels (curvature fp16 compared with tolerance);
  faithful mode (the reference's bit-fragile any(uv != MirrorUv(uv)) predicate): closed loop PSNR >= 45 dB on OUT_*."""
import os

import pytest
import torch

from nrd_sample_b200 import nrd_api as api, synth
from tests.util import compare

pytestmark = pytest.mark.gpu
F16 = api.Format.RGBA16_SFLOAT
RT = api.ResourceType
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reblur_96x64.pt")


@pytest.fixture(scope="module")
def ex():
    from nrd_sample_b200 import executor
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    executor.load()
    return executor


@pytest.fixture(scope="module")
def runner():
    from oracle import runner as r
    return r


def make_pair(ex, runner, w, h, robust):
    flags = ex.FLAG_QUAD_INTRINSICS | (ex.FLAG_ROBUST_MIRROR_TEST if robust else 0)
    cud = ex.CudaDenoiser(api.Denoiser.REBLUR_DIFFUSE_SPECULAR, w, h, flags=flags)
    orc = runner.OracleDenoiser(runner.default_host_library(), api.Denoiser.REBLUR_DIFFUSE_SPECULAR, w, h, robust_mirror_test=robust)
    g = {k: ex.alloc_texture(F16, w, h, "cuda:0") for k in ("d", "s")}
    c = {k: runner.alloc_texture(F16, w, h) for k in ("d", "s")}
    cud.set_user_texture(RT.OUT_DIFF_RADIANCE_HITDIST, g["d"], F16)
    cud.set_user_texture(RT.OUT_SPEC_RADIANCE_HITDIST, g["s"], F16)
    orc.set_user_texture(RT.OUT_DIFF_RADIANCE_HITDIST, c["d"])
    orc.set_user_texture(RT.OUT_SPEC_RADIANCE_HITDIST, c["s"])
    return cud, orc, g, c


def feed_both(cud, orc, runner, frame, keep):
    for k, v in frame.items():
        rt = getattr(RT, k)
        orc.set_user_texture(rt, v)
        keep[k] = v.to("cuda:0")
        cud.set_user_texture(rt, keep[k], runner.USER_FORMATS[rt])


@pytest.mark.parametrize("w,h", [(208, 120), (96, 64)])
def test_per_pass_parity_strict(ex, runner, w, h):
    """Every dispatch of 6 frames (31 on frame 0 incl. clears) replayed on the GPU from the oracle's own pre-dispatch textures."""
    orc = runner.OracleDenoiser(runner.default_host_library(), api.Denoiser.REBLUR_DIFFUSE_SPECULAR, w, h, robust_mirror_test=True)
    orc.set_user_texture(RT.OUT_DIFF_RADIANCE_HITDIST, runner.alloc_texture(F16, w, h))
    orc.set_user_texture(RT.OUT_SPEC_RADIANCE_HITDIST, runner.alloc_texture(F16, w, h))
    flags = ex.FLAG_QUAD_INTRINSICS | ex.FLAG_ROBUST_MIRROR_TEST
    worst = {}
    snap = {}

    def before(i, d, keys, den):
        snap["t"] = [den.textures[k].clone() for k in keys]

    def after(i, d, keys, den):
        gpu = [t.to("cuda:0") for t in snap["t"]]
        ex.dispatch(d.shader, d.constants, [ex.texture_of(g, den.formats[k]) for g, k in zip(gpu, keys)], flags=flags)
        torch.cuda.synchronize()
        for j, (b, k) in enumerate(zip(d.bindings, keys)):
            if b.descriptor != int(api.DescriptorType.STORAGE_TEXTURE):
                continue
            r = compare(gpu[j], den.textures[k], den.formats[k])
            key = (d.name, j, api.Format(den.formats[k]).name)
            if key not in worst or r["frac_bad"] > worst[key]["frac_bad"]:
                worst[key] = r

    for f in range(6):
        for k, v in synth.reblur_frame(f, w, h).items():
            orc.set_user_texture(getattr(RT, k), v)
        orc.denoise(synth.common_settings(f, w, h), before_dispatch=before, on_dispatch=after)

    assert len({k[0] for k in worst}) == 9  # 7 passes + 2 clear flavours
    for key, r in worst.items():
        # data2's fp16 curvature is chaotic on the static-camera frame 0 (0/0-like parallax direction, TA:387-467) and is ignored
        # there by the algorithm itself (no history yet); the occlusion bits / history amount / CatRom flag must still match
        limit = 3e-2 if key[2] == "R32_UINT" else 1e-3
        assert r["frac_bad"] <= limit, f"{key}: {r}"
        if key[2] in ("RGBA16_SFLOAT", "R32_SFLOAT") or (key[2] == "R16_SFLOAT" and "Pre-pass" not in key[0]):
            assert r["psnr"] >= 60.0, f"{key}: {r}"


def test_closed_loop_strict_and_faithful(ex, runner):
    w, h, n = 256, 144, 12
    # closed loop = recurrent system: per-pass differences of ~1 fp16 ulp feed back through the history for 12 frames
    for robust, min_psnr in ((True, 60.0), (False, 45.0)):
        cud, orc, g, c = make_pair(ex, runner, w, h, robust)
        keep = {}
        for f in range(n):
            feed_both(cud, orc, runner, synth.reblur_frame(f, w, h), keep)
            cs = synth.common_settings(f, w, h)
            orc.denoise(cs)
            cud.set_common_settings(cs)
            cud.denoise()
            torch.cuda.synchronize()
            for k in ("d", "s"):
                r = compare(g[k], c[k], F16)
                assert r["psnr"] >= min_psnr, f"robust={robust} frame {f} {k}: {r}"
                if robust:
                    assert r["frac_bad"] <= 2e-3, f"frame {f} {k}: {r}"
        cud.close()


def test_against_committed_golden_fixture(ex, runner):
    g = torch.load(GOLDEN)
    w, h = g["width"], g["height"]
    cud = ex.CudaDenoiser(api.Denoiser.REBLUR_DIFFUSE_SPECULAR, w, h, flags=ex.FLAG_QUAD_INTRINSICS | ex.FLAG_ROBUST_MIRROR_TEST)
    od, os_ = ex.alloc_texture(F16, w, h, "cuda:0"), ex.alloc_texture(F16, w, h, "cuda:0")
    cud.set_user_texture(RT.OUT_DIFF_RADIANCE_HITDIST, od, F16)
    cud.set_user_texture(RT.OUT_SPEC_RADIANCE_HITDIST, os_, F16)
    keep = {}
    for f, frame in enumerate(g["inputs"]):
        for k, v in frame.items():
            rt = getattr(RT, k)
            keep[k] = v.to("cuda:0")
            cud.set_user_texture(rt, keep[k], runner.USER_FORMATS[rt])
        cud.set_common_settings(synth.common_settings(f, w, h))
        cud.denoise()
        torch.cuda.synchronize()
        for got, want in ((od, g["strict"][f][0]), (os_, g["strict"][f][1])):
            r = compare(got, want, F16)
            assert r["frac_bad"] <= 2e-3 and r["psnr"] >= 70.0, f"frame {f}: {r}"
    cud.close()


def test_host_buffer_entry_point_matches_device_path(ex, runner):
    """nrdcuDenoiseHost (H2D + chain + D2H inside the call) == nrdcuDenoise on resident textures, bit for bit."""
    w, h = 160, 96
    a = ex.CudaDenoiser(api.Denoiser.REBLUR_DIFFUSE_SPECULAR, w, h)
    b = ex.CudaDenoiser(api.Denoiser.REBLUR_DIFFUSE_SPECULAR, w, h)
    od, os_ = ex.alloc_texture(F16, w, h, "cuda:0"), ex.alloc_texture(F16, w, h, "cuda:0")
    hd, hs = torch.zeros(h, w, 4, dtype=torch.float16).pin_memory(), torch.zeros(h, w, 4, dtype=torch.float16).pin_memory()
    a.set_user_texture(RT.OUT_DIFF_RADIANCE_HITDIST, od, F16)
    a.set_user_texture(RT.OUT_SPEC_RADIANCE_HITDIST, os_, F16)
    keep = {}
    for f in range(3):
        frame = synth.reblur_frame(f, w, h)
        host = {k: v.clone().pin_memory() for k, v in frame.items()}
        for k, v in frame.items():
            rt = getattr(RT, k)
            keep[k] = v.to("cuda:0")
            a.set_user_texture(rt, keep[k], runner.USER_FORMATS[rt])
            b.set_host_texture(rt, host[k], runner.USER_FORMATS[rt], is_output=False)
        b.set_host_texture(RT.OUT_DIFF_RADIANCE_HITDIST, hd, F16, is_output=True)
        b.set_host_texture(RT.OUT_SPEC_RADIANCE_HITDIST, hs, F16, is_output=True)
        cs = synth.common_settings(f, w, h)
        a.set_common_settings(cs)
        b.set_common_settings(cs)
        a.denoise()
        b.denoise_host()
        torch.cuda.synchronize()
        assert torch.equal(od.cpu().view(torch.int16), hd.view(torch.int16))
        assert torch.equal(os_.cpu().view(torch.int16), hs.view(torch.int16))
    a.close()
    b.close()


def test_full_size_invariants_1440p(ex):
    """At BASELINE.json's full size the oracle is too slow to ride along; check size-independent properties instead:
    finite, luma >= 0, normalised hit distance in [0,1], history length grows to the cap, determinism (same inputs -> same bits)."""
    w, h, n = 2560, 1440, 10
    outs = []
    for rep in range(2):
        cud = ex.CudaDenoiser(api.Denoiser.REBLUR_DIFFUSE_SPECULAR, w, h)
        od, os_ = ex.alloc_texture(F16, w, h, "cuda:0"), ex.alloc_texture(F16, w, h, "cuda:0")
        cud.set_user_texture(RT.OUT_DIFF_RADIANCE_HITDIST, od, F16)
        cud.set_user_texture(RT.OUT_SPEC_RADIANCE_HITDIST, os_, F16)
        fmts = {"IN_VIEWZ": api.Format.R32_SFLOAT, "IN_NORMAL_ROUGHNESS": api.Format.R10_G10_B10_A2_UNORM, "IN_MV": F16, "IN_DIFF_RADIANCE_HITDIST": F16,
                "IN_SPEC_RADIANCE_HITDIST": F16}
        for f in range(n):
            frame = synth.reblur_frame(f, w, h, device="cuda:0", with_clean=(f == n - 1))
            for k, v in frame.items():
                if not k.startswith("_"):
                    cud.set_user_texture(getattr(RT, k), v, fmts[k])
            cud.set_common_settings(synth.common_settings(f, w, h))
            cud.denoise()
            torch.cuda.synchronize()
        outs.append((od.clone(), os_.clone()))
        m = frame["_hit"]
        for out, clean_key, noisy_key in ((od, "_clean_diff", "IN_DIFF_RADIANCE_HITDIST"), (os_, "_clean_spec", "IN_SPEC_RADIANCE_HITDIST")):
            o = out.float()
            assert torch.isfinite(o).all()
            assert (o[..., 0] >= 0).all() and (o[..., 3] >= 0).all() and (o[..., 3] <= 1).all()
            tm = lambda x: x / (1 + x)  # noqa: E731
            clean = tm(frame[clean_key])
            mse_out = ((tm(synth.unpack_radiance(out)) - clean) ** 2)[m].mean()
            mse_in = ((tm(synth.unpack_radiance(frame[noisy_key])) - clean) ** 2)[m].mean()
            assert mse_out * 10 < mse_in, "denoised output must be >= 10 dB closer to the clean signal than the noisy input"
        idata = cud.pool_texture(True, 2).to(torch.int32) & 0xFFFF
        assert (idata & 63)[m].float().mean() > n - 3 and (idata & 63).max() <= 30
        cud.close()
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1]), "two runs on the same inputs must agree bit for bit"


def test_missing_resource_and_bad_format_are_reported(ex):
    cud = ex.CudaDenoiser(api.Denoiser.REBLUR_DIFFUSE_SPECULAR, 64, 64)
    cud.set_common_settings(synth.common_settings(0, 64, 64))
    with pytest.raises(ex.NrdcuError, match="INVALID_ARGUMENT"):
        cud.denoise()   # no user textures bound
    t = ex.alloc_texture(api.Format.R32_SFLOAT, 64, 64, "cuda:0")
    with pytest.raises(ex.NrdcuError):
        ex.dispatch("REBLUR_ClassifyTiles.cs.hlsl", b"\0" * 864, [ex.texture_of(t, api.Format.R32_SFLOAT), ex.texture_of(t, api.Format.R32_SFLOAT)])  # tiles must be R8
    with pytest.raises(ex.NrdcuError, match="UNSUPPORTED"):
        ex.dispatch("REBLUR_NoSuchPass.cs.hlsl", b"\0" * 864, [])   # a shader identifier without a kernel
    cud.close()


def test_two_strips_on_one_gpu_equal_the_whole_frame(ex, runner):
    """nrdcuDispatchRows + the halo plan of nrd_sample_b200/tiling.py, emulated on ONE GPU: two full-size texture sets, each pass
    computes rows [0, y) on set A and [y, H) on set B, rows a strip does not own are poisoned and only the 64-row halo is copied
    across the seam. The strips must reproduce the single-launch frame bit for bit (the multi-GPU run is tools/tiled_check.py)."""
    from nrd_sample_b200 import tiling
    w, h, frames = 256, 208, 5
    dev = "cuda:0"
    host = runner.default_host_library()
    strips = tiling.strip_rows(h, 2)
    inst = api.NrdInstance(host, [(0, api.Denoiser.REBLUR_DIFFUSE_SPECULAR)])
    perm, tran = inst.pools()

    def texture_set():
        t = {}
        for kind, pool in ((RT.PERMANENT_POOL, perm), (RT.TRANSIENT_POOL, tran)):
            for i, (fmt, ds) in enumerate(pool):
                t[(int(kind), i)] = (ex.alloc_texture(fmt, (w + ds - 1) // ds, (h + ds - 1) // ds, dev), fmt)
        for rt in (RT.OUT_DIFF_RADIANCE_HITDIST, RT.OUT_SPEC_RADIANCE_HITDIST):
            t[(int(rt), 0)] = (ex.alloc_texture(F16, w, h, dev), F16)
        return t

    sets = [texture_set(), texture_set(), texture_set()]   # strip A, strip B, whole frame
    pools = (int(RT.PERMANENT_POOL), int(RT.TRANSIENT_POOL))
    table = tiling.derive_halo_table(host, api.Denoiser.REBLUR_DIFFUSE_SPECULAR, w, h)   # per-texture aprons, derived from the readers
    for f in range(frames):
        frame = synth.reblur_frame(f, w, h)
        for s in sets:
            for k, v in frame.items():
                rt = getattr(RT, k)
                s[(int(rt), 0)] = (v.to(dev), runner.USER_FORMATS[rt])
        assert inst.set_common_settings(synth.common_settings(f, w, h)) == api.Result.SUCCESS
        r, dispatches = inst.get_compute_dispatches([0])
        assert r == api.Result.SUCCESS
        for d in dispatches:
            keys = [(b.type, b.index) if b.type in pools else (b.type, 0) for b in d.bindings]
            for si, rows in ((0, strips[0]), (1, strips[1]), (2, None)):
                tex = [ex.texture_of(*sets[si][k]) for k in keys]
                ex.dispatch(d.shader, d.constants, tex, flags=ex.FLAG_QUAD_INTRINSICS | ex.FLAG_ROBUST_MIRROR_TEST, rows=rows)
            if d.shader.startswith("Clear") or d.name.endswith("Classify tiles"):   # clears cover whole textures; the tile mask is strip-local
                continue
            planes, halos = [[], []], []
            for j, (b, k) in enumerate(zip(d.bindings, keys)):
                if b.descriptor != int(api.DescriptorType.STORAGE_TEXTURE):
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
        for rt in (RT.OUT_DIFF_RADIANCE_HITDIST, RT.OUT_SPEC_RADIANCE_HITDIST):
            whole = sets[2][(int(rt), 0)][0]
            for si, (y0, y1) in enumerate(strips):
                got = sets[si][(int(rt), 0)][0]
                assert torch.equal(got[y0:y1].view(torch.int16), whole[y0:y1].view(torch.int16)), f"frame {f} strip {si} {rt}"


def test_pipelined_host_path_matches_device_path(ex, runner):
    """nrdcuDenoiseHostPipelined (uploads / downloads of neighbouring frames overlapping the kernels, double-buffered inputs) must hand
    back, frame by frame, exactly the bits of nrdcuDenoise on resident textures."""
    w, h, n = 320, 192, 6
    a = ex.CudaDenoiser(api.Denoiser.REBLUR_DIFFUSE_SPECULAR, w, h)
    b = ex.CudaDenoiser(api.Denoiser.REBLUR_DIFFUSE_SPECULAR, w, h)
    od, os_ = ex.alloc_texture(F16, w, h, "cuda:0"), ex.alloc_texture(F16, w, h, "cuda:0")
    a.set_user_texture(RT.OUT_DIFF_RADIANCE_HITDIST, od, F16)
    a.set_user_texture(RT.OUT_SPEC_RADIANCE_HITDIST, os_, F16)
    want, got, keep_host = [], [], []
    keep = {}
    for f in range(n):
        frame = synth.reblur_frame(f, w, h)
        host = {k: v.clone().pin_memory() for k, v in frame.items()}
        hd, hs = torch.zeros(h, w, 4, dtype=torch.float16).pin_memory(), torch.zeros(h, w, 4, dtype=torch.float16).pin_memory()
        keep_host.append(host)
        for k, v in frame.items():
            rt = getattr(RT, k)
            keep[k] = v.to("cuda:0")
            a.set_user_texture(rt, keep[k], runner.USER_FORMATS[rt])
            b.set_host_texture(rt, host[k], runner.USER_FORMATS[rt], is_output=False)
        b.set_host_texture(RT.OUT_DIFF_RADIANCE_HITDIST, hd, F16, is_output=True)
        b.set_host_texture(RT.OUT_SPEC_RADIANCE_HITDIST, hs, F16, is_output=True)
        cs = synth.common_settings(f, w, h)
        a.set_common_settings(cs)
        b.set_common_settings(cs)
        a.denoise()
        b.denoise_host_pipelined()   # no synchronisation between frames: copies of frame f + 1 overlap the kernels of frame f
        torch.cuda.current_stream().synchronize()   # stream only: the copy streams of `b` keep running
        want.append((od.cpu().clone(), os_.cpu().clone()))
        got.append((hd, hs))
    b.host_flush()
    torch.cuda.synchronize()
    for f in range(n):
        assert torch.equal(want[f][0].view(torch.int16), got[f][0].view(torch.int16)), f"frame {f} diffuse"
        assert torch.equal(want[f][1].view(torch.int16), got[f][1].view(torch.int16)), f"frame {f} specular"
    a.close()
    b.close()


@pytest.mark.parametrize("mode", [1, 2])
def test_hit_distance_reconstruction_parity(ex, runner, mode):
    """ReblurSettings::hitDistanceReconstructionMode = AREA_3X3 / AREA_5X5 (REBLUR_HitDistReconstruction.cs.hlsl, the NRD README's
    benchmark setting) on inputs where every pixel traced one lobe only: the pass itself replayed from the oracle's pre-dispatch
    textures, then the closed loop over 6 frames."""
    w, h = 208, 120
    settings = api.ReblurSettings(hitDistanceReconstructionMode=mode)
    cud, orc, g, c = make_pair(ex, runner, w, h, True)
    flags = ex.FLAG_QUAD_INTRINSICS | ex.FLAG_ROBUST_MIRROR_TEST
    snap, seen = {}, []

    def before(i, d, keys, den):
        if "HitDistReconstruction" in d.shader:
            snap["t"] = [den.textures[k].clone() for k in keys]

    def after(i, d, keys, den):
        if "HitDistReconstruction" not in d.shader:
            return
        assert d.shader.endswith(f"MODE_5X5={mode - 1}")
        gpu = [t.to("cuda:0") for t in snap["t"]]
        ex.dispatch(d.shader, d.constants, [ex.texture_of(t, den.formats[k]) for t, k in zip(gpu, keys)], flags=flags)
        torch.cuda.synchronize()
        for j in (5, 6):
            r = compare(gpu[j], den.textures[keys[j]], den.formats[keys[j]])
            assert r["frac_bad"] <= 1e-4 and r["psnr"] >= 70.0, f"lobe {j - 5}: {r}"
            seen.append(r)

    keep = {}
    cud.set_denoiser_settings(settings)
    for f in range(6):
        feed_both(cud, orc, runner, synth.reblur_frame(f, w, h, holes=True), keep)
        cs = synth.common_settings(f, w, h)
        orc.denoise(cs, settings=settings, before_dispatch=before, on_dispatch=after)
        cud.set_common_settings(cs)
        cud.denoise()
        torch.cuda.synchronize()
        for k in ("d", "s"):
            r = compare(g[k], c[k], F16)
            assert r["psnr"] >= 60.0 and r["frac_bad"] <= 2e-3, f"mode {mode} frame {f} {k}: {r}"
    assert len(seen) == 12
    cud.close()
