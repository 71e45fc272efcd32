"""Multi-GPU strip tiling (nrd_sample_b200/tiling.py, BASELINE.json config 3) — host-side logic on CPU with gloo.

The CPU oracle stands in for the kernels: every rank replays each dispatch on its own full-size texture set, then
POISONS every row of the written textures outside its strip (so nothing but the halo exchange can provide them), then
trades halos through `tiling.exchange_halos` over a world_size-2 / -3 gloo group. The strips of the final outputs must
equal the single-process frame bit for bit — which holds only if the halo covers the reach of every later pass."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nrd_sample_b200 import nrd_api as api, synth, tiling

RT = api.ResourceType
F16 = api.Format.RGBA16_SFLOAT


def test_strip_rows():
    assert tiling.strip_rows(2160, 8) == [(0, 272), (272, 544), (544, 816), (816, 1088), (1088, 1360), (1360, 1632), (1632, 1904), (1904, 2160)]
    assert tiling.strip_rows(2160, 2) == [(0, 1088), (1088, 2160)]
    for h, n in ((1440, 4), (1080, 8), (200, 3), (64, 4)):
        s = tiling.strip_rows(h, n)
        assert s[0][0] == 0 and s[-1][1] == h and all(a[1] == b[0] for a, b in zip(s, s[1:]))
        assert all(a % 16 == 0 for a, _ in s) and all(b > a for a, b in s)
    with pytest.raises(ValueError):
        tiling.strip_rows(32, 4)


def test_balanced_strips_follow_the_work():
    """Sky rows cost nothing: cuts by cumulative denoising-range pixels give every rank the same work, not the same height."""
    h = 544
    w = tiling.tile_row_weights(synth.reblur_frame(0, 320, h)["IN_VIEWZ"])
    assert len(w) == h // 16 and w[0] < 0.1 < w[-2]   # sky on top, geometry below
    for world in (2, 4):
        s = tiling.strip_rows(h, world, w, min_rows=64)
        assert s[0][0] == 0 and s[-1][1] == h and all(a[1] == b[0] for a, b in zip(s, s[1:])) and all(a % 16 == 0 for a, _ in s)
        assert all(b - a >= 64 for a, b in s)
        work = [sum(w[a // 16:(b + 15) // 16]) for a, b in s]
        even = [sum(w[a // 16:(b + 15) // 16]) for a, b in tiling.strip_rows(h, world)]
        assert max(work) < max(even) * 0.85 and max(work) / (sum(work) / world) < 1.2


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


# denoiser -> ( outputs, the other lobe's input marker to drop, SH inputs, CPU engine ): the hand-written oracle covers the RADIANCE denoisers; the
# SH one runs on the reference's own shaders ( oracle/_ref/libnrd_refshaders.so, skipped when it was not built )
VARIANTS = {
    "REBLUR_DIFFUSE_SPECULAR": (("OUT_DIFF_RADIANCE_HITDIST", "OUT_SPEC_RADIANCE_HITDIST"), "~", False, "oracle"),
    "REBLUR_DIFFUSE": (("OUT_DIFF_RADIANCE_HITDIST",), "_SPEC_", False, "oracle"),
    "REBLUR_SPECULAR": (("OUT_SPEC_RADIANCE_HITDIST",), "_DIFF_", False, "oracle"),
    "REBLUR_DIFFUSE_SPECULAR_SH": (("OUT_DIFF_SH0", "OUT_DIFF_SH1", "OUT_SPEC_SH0", "OUT_SPEC_SH1"), "~", True, "reference"),
}


def _frame(name, f, w, h):
    _, drop, sh, _ = VARIANTS[name]
    return {k: v for k, v in synth.reblur_frame(f, w, h, sh=sh).items() if drop not in k}


def _make(name, w, h):
    from oracle import runner
    outputs, _, _, engine = VARIANTS[name]
    den = runner.OracleDenoiser(runner.default_host_library(), getattr(api.Denoiser, name), w, h, robust_mirror_test=True, engine=engine)
    outs = [runner.alloc_texture(F16, w, h) for _ in outputs]
    for o, t in zip(outputs, outs):
        den.set_user_texture(getattr(RT, o), t)
    return den, outs


def _reference_run(w, h, frames, name="REBLUR_DIFFUSE_SPECULAR"):
    den, outs = _make(name, w, h)
    res = []
    for f in range(frames):
        for k, v in _frame(name, f, w, h).items():
            den.set_user_texture(getattr(RT, k), v)
        den.denoise(synth.common_settings(f, w, h))
        res.append(tuple(t.clone() for t in outs))
    return res


def _rank_main(rank, world, port, w, h, frames, halo, result_dir, name="REBLUR_DIFFUSE_SPECULAR"):
    from oracle import runner
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        runner.lib().nrd_oracle_set_threads(2)
        strips = tiling.strip_rows(h, world)
        y0, y1 = strips[rank]
        den, outs_t = _make(name, w, h)
        table = tiling.derive_halo_table(runner.default_host_library(), getattr(api.Denoiser, name), w, h, halo)   # per-denoiser aprons ( tiling.py )

        def after(i, d, keys, self):
            if d.shader.startswith("Clear"):
                return  # every rank clears its whole copy
            if d.name.endswith("Classify tiles"):
                return  # the tile mask is strip-local ( read at pixel >> 4 ); REBLUR_DIFFUSE_SPECULAR_SH keeps it in a full-resolution texture whose rows do not map to strips
            planes, halos = [], []
            for j, (b, k) in enumerate(zip(d.bindings, keys)):
                if b.descriptor != int(api.DescriptorType.STORAGE_TEXTURE):
                    continue
                t = self.textures[k]
                p = t.view(torch.uint8).view(t.shape[0], -1)
                ty0, ty1, _ = tiling._scaled((y0, y1), halo, p.shape[0], h)
                if k[0] != int(RT.IN_MV):   # bound read-write by temporal stabilisation but only written on clear frames: stays an input
                    p[:ty0] = 0xFF          # poison: 0xFFFF is a NaN in fp16, 255 in UNORM, an impossible history word
                    p[ty1:] = 0xFF
                planes.append(p)
                halos.append(tiling.halo_rows_for(table, d.name, j, halo))   # the derived per-texture table, capped by `halo`
            tiling.exchange_halos(planes, strips, rank, h, halos)

        outs = []
        for f in range(frames):
            for k, v in _frame(name, f, w, h).items():
                den.set_user_texture(getattr(RT, k), v)
            den.denoise(synth.common_settings(f, w, h), on_dispatch=after)
            outs.append(tuple(t[y0:y1].clone() for t in outs_t))
        torch.save(outs, os.path.join(result_dir, f"rank{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,h,halo,expect_equal", [(2, 160, 64, True), (3, 208, 64, True), (2, 160, 16, False)])
def test_strips_with_halo_exchange_reproduce_the_single_process_frame(tmp_path, world, h, halo, expect_equal):
    w, frames = 96, 4
    ref = _reference_run(w, h, frames)
    mp.spawn(_rank_main, args=(world, _free_port(), w, h, frames, halo, str(tmp_path)), nprocs=world, join=True)
    strips = tiling.strip_rows(h, world)
    equal = True
    for r, (y0, y1) in enumerate(strips):
        outs = torch.load(os.path.join(str(tmp_path), f"rank{r}.pt"))
        for f in range(frames):
            for got, want in zip(outs[f], ref[f]):
                same = torch.equal(got.view(torch.int16), want[y0:y1].view(torch.int16))
                if expect_equal:
                    assert same, f"world {world} rank {r} frame {f}: strip differs from the single-process frame"
                equal &= same
    if not expect_equal:
        assert not equal, "a 16-row halo cannot cover 30-60 px blur radii: the poison must have leaked into the strips"


def test_halo_table_is_derived_per_denoiser():
    """The apron of a written texture follows its READERS, not a binding index: the same role gets the same rows in every REBLUR variant."""
    from oracle import runner
    lib = runner.default_host_library()
    rows = {}
    for name in ("REBLUR_DIFFUSE_SPECULAR", "REBLUR_DIFFUSE", "REBLUR_SPECULAR", "REBLUR_DIFFUSE_SH", "REBLUR_SPECULAR_SH", "REBLUR_DIFFUSE_SPECULAR_SH"):
        t = tiling.derive_halo_table(lib, getattr(api.Denoiser, name), 96, 160)
        assert t and all(0 <= v <= tiling.HALO_ROWS for v in t.values())
        by_pass = {}
        for (p, _), v in t.items():
            by_pass.setdefault(p, []).append(v)
        rows[name] = by_pass
        assert 32 in by_pass["Temporal accumulation"] and 0 in by_pass["Temporal accumulation"]          # radiance -> history fix taps; data2 -> at the pixel
        assert by_pass["Classify tiles"] == [0]
        assert max(by_pass["Blur"]) == 64                                                                # post-blur reaches 2 x 30 rows; PREV_VIEWZ survives the frame
        assert 32 in by_pass["History fix"] and max(by_pass["Temporal stabilization"]) == 64
        assert 0 in by_pass["Temporal stabilization"]                                                    # the final outputs / IN_MV
    # larger radii -> taller aprons
    big = tiling.derive_halo_table(lib, api.Denoiser.REBLUR_DIFFUSE_SPECULAR, 96, 160, 128, api.ReblurSettings(maxBlurRadius=50.0))
    assert max(v for (p, _), v in big.items() if p == "Blur") >= 102


@pytest.mark.parametrize("name", ["REBLUR_DIFFUSE", "REBLUR_SPECULAR", "REBLUR_DIFFUSE_SPECULAR_SH"])
def test_strips_of_the_other_reblur_denoisers(tmp_path, name):
    """the binding-index table only fitted REBLUR_DIFFUSE_SPECULAR; the derived table must hold for the other variants."""
    from oracle import runner
    if VARIANTS[name][3] == "reference" and runner.ref_shaders() is None:
        pytest.skip("oracle/_ref/libnrd_refshaders.so was not shipped")
    world, w, h, frames, halo = 2, 96, 160, 3, 64
    ref = _reference_run(w, h, frames, name)
    mp.spawn(_rank_main, args=(world, _free_port(), w, h, frames, halo, str(tmp_path), name), nprocs=world, join=True)
    for r, (y0, y1) in enumerate(tiling.strip_rows(h, world)):
        outs = torch.load(os.path.join(str(tmp_path), f"rank{r}.pt"))
        for f in range(frames):
            for got, want in zip(outs[f], ref[f]):
                assert torch.equal(got.view(torch.int16), want[y0:y1].view(torch.int16)), f"{name} rank {r} frame {f}: strip differs from the single-process frame"


def test_sigma_and_relax_get_the_full_apron():
    """REBLUR's per-pass reach rules must not leak onto SIGMA / RELAX passes of the same name ( SIGMA's temporal stabilization reads a history copy written the same
    frame at pixel + motion; its history length is R32_UINT like REBLUR's at-the-pixel data2 ): every full-resolution texture they write gets the default apron."""
    from oracle import runner
    lib = runner.default_host_library()
    for name in ("SIGMA_SHADOW", "SIGMA_SHADOW_TRANSLUCENCY", "RELAX_DIFFUSE_SPECULAR_SH", "RELAX_DIFFUSE"):
        t = tiling.derive_halo_table(lib, getattr(api.Denoiser, name), 96, 160)
        full = {k: v for k, v in t.items() if k[0] not in ("Classify tiles", "Smooth tiles")}
        assert full and all(v in (0, tiling.HALO_ROWS) for v in full.values()), (name, t)      # 0: written, never read again before the next write ( final outputs )
        assert sum(1 for v in full.values() if v == tiling.HALO_ROWS) >= 4, (name, t)
        assert all(v == 0 for k, v in t.items() if k[0] in ("Classify tiles", "Smooth tiles")), (name, t)
