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


def _reference_run(w, h, frames):
    from oracle import runner
    den = runner.OracleDenoiser(runner.default_host_library(), api.Denoiser.REBLUR_DIFFUSE_SPECULAR, w, h, robust_mirror_test=True)
    od, os_ = runner.alloc_texture(F16, w, h), runner.alloc_texture(F16, w, h)
    den.set_user_texture(RT.OUT_DIFF_RADIANCE_HITDIST, od)
    den.set_user_texture(RT.OUT_SPEC_RADIANCE_HITDIST, os_)
    outs = []
    for f in range(frames):
        for k, v in synth.reblur_frame(f, w, h).items():
            den.set_user_texture(getattr(RT, k), v)
        den.denoise(synth.common_settings(f, w, h))
        outs.append((od.clone(), os_.clone()))
    return outs


def _rank_main(rank, world, port, w, h, frames, halo, result_dir):
    from oracle import runner
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        runner.lib().nrd_oracle_set_threads(2)
        strips = tiling.strip_rows(h, world)
        y0, y1 = strips[rank]
        den = runner.OracleDenoiser(runner.default_host_library(), api.Denoiser.REBLUR_DIFFUSE_SPECULAR, w, h, robust_mirror_test=True)
        od, os_ = runner.alloc_texture(F16, w, h), runner.alloc_texture(F16, w, h)
        den.set_user_texture(RT.OUT_DIFF_RADIANCE_HITDIST, od)
        den.set_user_texture(RT.OUT_SPEC_RADIANCE_HITDIST, os_)

        def after(i, d, keys, self):
            if d.shader.startswith("Clear"):
                return  # every rank clears its whole copy
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
                halos.append(tiling.halo_rows_for(d.name, j, halo))   # the per-texture table of tiling.py, capped by `halo`
            tiling.exchange_halos(planes, strips, rank, h, halos)

        outs = []
        for f in range(frames):
            for k, v in synth.reblur_frame(f, w, h).items():
                den.set_user_texture(getattr(RT, k), v)
            den.denoise(synth.common_settings(f, w, h), on_dispatch=after)
            outs.append((od[y0:y1].clone(), os_[y0:y1].clone()))
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
