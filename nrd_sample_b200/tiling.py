"""One frame tiled over N GPUs as horizontal strips with halo exchange at the seams
(BASELINE.json config 3: REBLUR_DIFFUSE_SPECULAR 3840x2160 over 2/4/8 B200).

Layout. Every rank owns a full-resolution nrdcuContext (pools are 0.6 GB at 4K — memory is not the constraint) but
computes only the rows [y0, y1) of its strip in every pass (`nrdcuDenoiseRows`). Strip boundaries are multiples of 16
rows, so the 16x16 sky-tile classification and every CTA (32x8 / 32x16 pixels) are strip-local. The noisy inputs are
provided full frame (a tiled renderer would render its strip plus the apron).

Exchange. A pass reads its inputs through gathers that reach up to 60 rows away (post-blur: 2 x maxBlurRadius;
history fix: 2 x 14 + 4; temporal passes: motion vector + 2), and what it reads was written by an earlier pass — or by
the previous frame — on the NEIGHBOUR rank for rows outside the strip. So after every dispatch each rank sends the
`halo` rows at the edges of its strip of every texture the dispatch wrote to the neighbour above / below, and receives
the neighbour's rows into the apron of its own copy: one batched NCCL send/recv group per pass over NVLink, 6-7 per
frame. With halo >= the largest reach the N-GPU result is bit-identical to the single-GPU frame (tests/test_tiling.py
checks exactly that, on gloo with the CPU oracle standing in for the kernels and on NCCL with the real ones).

Two transports. `mode="peer"` (default on GPUs): neighbouring ranks map each other's textures through CUDA IPC and the
executor itself stores the seam rows into the neighbours' copies over NVLink after every pass, then raises a flag the
neighbour spins on (csrc/kernels/peer_halo.cu): three tiny launches per pass, nothing on the host. `mode="nccl"`: the
same rows through torch.distributed batched send/recv from a per-dispatch callback — the portable baseline (also what
the gloo CPU tests exercise), ~0.1 ms slower per pass.

Bound on temporal reach: history is fetched at pixel + motion; HALO_ROWS - 2 = 62 rows of vertical motion per frame
are covered, beyond that a strip needs a taller halo (`halo_rows=`). The bound is CHECKED: every strip frame the executor
scans the strip's motion vectors on the device and keeps the worst overshoot (`TiledDenoiser.motion_bound()`).
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

TILE = 16
HALO_ROWS = 64

# Rows of apron each written texture needs on the neighbour: derived per denoiser from WHO READS the texture later (this frame or the
# next one) and how far that reader reaches from its own pixel. Keyed by the reader's pass name ( "<denoiser> - <pass>" -> "<pass>" ) and the
# format class of the texture it reads:
#   0  = only ever read at the pixel itself,  2 = read through a 1-texel shared-memory border,
#   history fix: 5x5 taps with stride <= historyFixBasePixelStride (2 x 14 + 4 = 32, data1 rides along with the taps),
#   blur: maxBlurRadius + 2, post-blur: 2 x maxBlurRadius + 2, pre-pass: the larger pre-pass radius + 2,
#   anything read by the NEXT frame is fetched at pixel + motion: the default.
_DATA2 = ("R32_UINT", "R8_UINT")
_DATA1 = ("RG8_UNORM", "R8_UNORM")


def reader_reach(pass_name: str, fmt_name: str, same_frame: bool, default: int, max_blur_radius: float = 30.0, prepass_radius: float = 50.0, history_fix_stride: float = 14.0,
                 reblur: bool = True) -> int:
    """Rows a REBLUR pass reaches beyond its own pixel into an input texture of format `fmt_name` written `same_frame` or last frame."""
    if not same_frame:
        return default
    # the per-pass rules below are REBLUR's. SIGMA and RELAX share some pass NAMES with other access patterns ( SIGMA's temporal stabilization reads a history copy
    # written the same frame, at pixel + motion ): every texture of theirs gets the full apron — their reaches ( SIGMA blur 32 + 2, RELAX a-trous strides <= 16 + 1,
    # pre-pass 50 + 2 ) all fit, and the extra seam bytes are small against REBLUR's
    # ( decided by the denoiser, not by the name's prefix: REBLUR_DIFFUSE_SH's passes are called "DENOISER_NAME - ...", a wart of the reference kept as is )
    if not reblur:
        return default
    name = pass_name.split(" - ")[-1]
    data = fmt_name in _DATA1 or fmt_name in _DATA2
    if name == "Pre-pass":
        r = int(prepass_radius + 0.999) + 2
    elif name == "Temporal accumulation":
        r = 0 if data else 2
    elif name == "History fix":
        r = 0 if fmt_name in _DATA2 else int(2.0 * history_fix_stride + 0.999) + 4
    elif name == "Blur":
        r = 0 if data else int(max_blur_radius + 0.999) + 2
    elif name == "Post-blur":
        r = 0 if data else int(2.0 * max_blur_radius + 0.999) + 2
    elif name == "Temporal stabilization":
        r = 0 if data else 2
    elif name in ("Split screen", "Classify tiles"):
        r = 0
    else:
        r = default
    return min(default, r)


def derive_halo_table(host_lib, denoiser: int, width: int, height: int, default: int = HALO_ROWS, settings=None, common_kw=None) -> dict:
    """{(pass name, binding index): apron rows} for every storage binding of a steady-state frame of `denoiser`, from the dispatch streams
    of two consecutive frames ( the pools ping-pong with period 2 ) of a scratch nrd::Instance: a texture written by dispatch i needs as many
    rows on the neighbour as the farthest-reaching dispatch that reads it before it is written again. Works for every REBLUR denoiser
    ( binding indices differ between REBLUR_DIFFUSE / _SPECULAR / _DIFFUSE_SPECULAR and their SH variants; the readers do not )."""
    from . import nrd_api as api, synth
    inst = api.NrdInstance(host_lib, [(0, denoiser)])
    assert inst.result == api.Result.SUCCESS, inst.result
    perm, tran = inst.pools()
    fmt_of_pool = {int(api.ResourceType.PERMANENT_POOL): perm, int(api.ResourceType.TRANSIENT_POOL): tran}
    kw = {"reblur": int(denoiser) <= int(api.Denoiser.REBLUR_DIFFUSE_DIRECTIONAL_OCCLUSION)}
    if settings is not None:
        assert inst.set_denoiser_settings(0, settings) == api.Result.SUCCESS
        for src, dst in (("maxBlurRadius", "max_blur_radius"), ("historyFixBasePixelStride", "history_fix_stride")):
            if hasattr(settings, src):
                kw[dst] = float(getattr(settings, src))
        if hasattr(settings, "diffusePrepassBlurRadius"):
            kw["prepass_radius"] = max(float(settings.diffusePrepassBlurRadius), float(settings.specularPrepassBlurRadius))
    streams = []
    for f in range(4):
        assert inst.set_common_settings(synth.common_settings(f, width, height, **(common_kw or {}))) == api.Result.SUCCESS
        r, d = inst.get_compute_dispatches([0])
        assert r == api.Result.SUCCESS
        streams.append(d)
    inst.destroy()

    def key(b):
        return (b.type, b.index if b.type in fmt_of_pool else 0)

    def fmt_name(b):
        if b.type in fmt_of_pool:
            return api.Format(fmt_of_pool[b.type][b.index][0]).name
        return "RGBA16_SFLOAT"   # user outputs ( radiance / SH ): the widest class

    storage = int(api.DescriptorType.STORAGE_TEXTURE)
    table = {}
    for a, nxt in ((streams[2], streams[3]), (streams[3], streams[2])):   # steady state: no clears, both ping-pong phases
        seq = [(d, True) for d in a] + [(d, False) for d in nxt]
        for i, d in enumerate(a):
            for j, b in enumerate(d.bindings):
                if b.descriptor != storage:
                    continue
                rows = 0
                # ( REBLUR_DIFFUSE_SPECULAR_SH keeps the mask in a full-resolution texture, DESIGN.md "reference warts": same access pattern )
                if d.name.endswith("Classify tiles") or (b.type in fmt_of_pool and fmt_of_pool[b.type][b.index][1] > 1):
                    table[(d.name.split(" - ")[-1], j)] = 0   # the 1/16-resolution tile mask: read at ( pixel >> 4 ) by the strip that classified it
                    continue
                # IN_MV is bound read-write by temporal stabilisation but only written on clear frames: it stays a ( full-frame ) input
                for r, same in (seq[i + 1:] if b.type != int(api.ResourceType.IN_MV) else []):
                    rd, wr = False, False
                    for x in r.bindings:
                        if key(x) == key(b):
                            if x.descriptor == storage:
                                wr = True
                            else:
                                rd = True
                    if rd:
                        rows = max(rows, reader_reach(r.name, fmt_name(b), same, default, **kw))
                    if wr:
                        break
                k = (d.name.split(" - ")[-1], j)
                table[k] = max(table.get(k, 0), rows)
    return table


def halo_rows_for(table: dict, pass_name: str, binding: int, default: int = HALO_ROWS) -> int:
    """Apron rows for output `binding` of the pass called '<denoiser> - <pass_name>' (never more than `default`)."""
    return min(default, table.get((pass_name.split(" - ")[-1], binding), default))


def strip_rows(height: int, world: int, weights: Optional[Sequence[float]] = None, min_rows: int = 0) -> List[Tuple[int, int]]:
    """Rows [y0, y1) of each rank's strip: whole 16-row tile rows, the last strip takes the ragged end.

    Without `weights` the tile rows are split evenly. With `weights` (one cost per 16-row tile row, e.g. the number of pixels in
    denoising range — sky costs nothing, see `tile_row_weights`) the cuts sit where the cumulative cost crosses k / world of the
    total, so that every GPU gets the same amount of work rather than the same number of rows; `min_rows` keeps every strip at
    least as tall as the halo its neighbours read from it."""
    tiles = (height + TILE - 1) // TILE
    if world > tiles:
        raise ValueError(f"{world} strips need at least {world} tile rows, the frame has {tiles}")
    min_tiles = max(1, (min_rows + TILE - 1) // TILE)
    if weights is None:
        base, extra = divmod(tiles, world)
        counts = [base + (1 if r < extra else 0) for r in range(world)]
    else:
        w = [float(x) for x in weights]
        if len(w) != tiles:
            raise ValueError(f"expected {tiles} tile-row weights, got {len(w)}")
        total = sum(w) or 1.0
        cuts, acc, t = [0], 0.0, 0
        for k in range(1, world):
            target = total * k / world
            while t < tiles and acc + w[t] * 0.5 < target:
                acc += w[t]
                t += 1
            t = min(max(t, cuts[-1] + min_tiles), tiles - (world - k) * min_tiles)   # leave room for the strips still to come
            acc = sum(w[:t])
            cuts.append(t)
        cuts.append(tiles)
        counts = [b - a for a, b in zip(cuts, cuts[1:])]
    if min(counts) < 1:
        raise ValueError("a strip came out empty")
    out, t = [], 0
    for n in counts:
        out.append((t * TILE, min((t + n) * TILE, height)))
        t += n
    return out


def tile_row_weights(viewz: torch.Tensor, denoising_range: float = 5e5, floor: float = 0.03) -> List[float]:
    """Cost of each 16-row tile row of a frame: the fraction of its pixels inside the denoising range (every pass early-outs on sky
    tiles / pixels), plus a small floor for the fixed per-row cost. Every rank holds the full IN_VIEWZ, so all ranks derive the same cuts."""
    h = viewz.shape[0]
    inside = (viewz.abs() < denoising_range).float().mean(dim=1)
    pad = (-h) % TILE
    if pad:
        inside = torch.cat([inside, inside.new_zeros(pad)])
    return (inside.view(-1, TILE).mean(dim=1) + floor).tolist()


def _scaled(rows: Tuple[int, int], halo: int, tex_height: int, full_height: int) -> Tuple[int, int, int]:
    """Strip rows and halo in the row units of a texture that may be downsampled (tiles: 1/16)."""
    if tex_height == full_height:
        return rows[0], rows[1], halo
    ds = TILE   # the only downsampled textures are the 1/16 tile masks ( re-deriving the factor from the two heights is wrong for e.g. 130 rows -> 9 tile rows )
    assert tex_height == (full_height + TILE - 1) // TILE, (tex_height, full_height)
    return rows[0] // ds, min((rows[1] + ds - 1) // ds, tex_height), (halo + ds - 1) // ds


def build_exchange(planes: Sequence[torch.Tensor], strips: Sequence[Tuple[int, int]], rank: int, full_height: int, halo=HALO_ROWS,
                   group: Optional[dist.ProcessGroup] = None):
    """The send / recv list of one halo trade: ([P2POp], bytes sent). `halo` is one row count or one per plane."""
    world = len(strips)
    ops, sent = [], 0
    halos = [halo] * len(planes) if isinstance(halo, int) else list(halo)
    for p, hp in zip(planes, halos):
        assert p.dim() == 2 and p.is_contiguous()
        if world == 1 or hp <= 0:
            continue
        y0, y1, h = _scaled(strips[rank], hp, p.shape[0], full_height)
        for nb, send_rows, recv_rows in ((rank - 1, (y0, min(y0 + h, y1)), (max(y0 - h, 0), y0)), (rank + 1, (max(y1 - h, y0), y1), (y1, min(y1 + h, p.shape[0])))):
            if nb < 0 or nb >= world:
                continue
            # the neighbour's strip must be at least as tall as the halo it supplies, or rows would come from two ranks away
            ny0, ny1, _ = _scaled(strips[nb], hp, p.shape[0], full_height)
            assert ny1 - ny0 >= recv_rows[1] - recv_rows[0] and y1 - y0 >= send_rows[1] - send_rows[0], "strip shorter than the halo"
            if send_rows[1] > send_rows[0]:
                s = p[send_rows[0]:send_rows[1]]
                ops.append(dist.P2POp(dist.isend, s, nb, group))
                sent += s.numel() * s.element_size()
            if recv_rows[1] > recv_rows[0]:
                ops.append(dist.P2POp(dist.irecv, p[recv_rows[0]:recv_rows[1]], nb, group))
    return ops, sent


def exchange_halos(planes: Sequence[torch.Tensor], strips: Sequence[Tuple[int, int]], rank: int, full_height: int, halo=HALO_ROWS,
                   group: Optional[dist.ProcessGroup] = None) -> int:
    """Send the edge rows of this rank's strip of every plane to the neighbours and receive theirs into the apron.

    `planes`: 2-D row-major views [rows, row_bytes_or_elems] of whole textures (row slices are contiguous, so every
    message is one contiguous block). Works on CUDA tensors (NCCL) and CPU tensors (gloo). Returns the bytes sent."""
    ops, sent = build_exchange(planes, strips, rank, full_height, halo, group)
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return sent


def exchange_halos_local(plane_sets: Sequence[Sequence[torch.Tensor]], strips: Sequence[Tuple[int, int]], full_height: int, halo=HALO_ROWS) -> None:
    """The same row trade between strips that live in ONE process (plane_sets[r][i] = plane i of strip r): plain copies.
    Used to run several strips on a single GPU (tests, debugging) — the row arithmetic is shared with `exchange_halos`."""
    world = len(strips)
    halos = [halo] * len(plane_sets[0]) if isinstance(halo, int) else list(halo)
    for r in range(world - 1):
        for up, down, hp in zip(plane_sets[r], plane_sets[r + 1], halos):
            if hp <= 0:
                continue
            _, seam, h = _scaled(strips[r], hp, up.shape[0], full_height)   # seam = first row of strip r + 1
            u0, _, _ = _scaled(strips[r], hp, up.shape[0], full_height)
            _, d1, _ = _scaled(strips[r + 1], hp, up.shape[0], full_height)
            down[max(seam - h, u0):seam] = up[max(seam - h, u0):seam]        # bottom rows of strip r -> apron above strip r + 1
            up[seam:min(seam + h, d1)] = down[seam:min(seam + h, d1)]        # top rows of strip r + 1 -> apron below strip r


class TiledDenoiser:
    """N ranks, one frame: rank r denoises the rows of strip r and trades seam rows with its neighbours after every pass.
    Construct on every rank of an initialised process group (backend nccl); call `denoise()` in lockstep."""

    def __init__(self, denoiser: int, width: int, height: int, rank: int, world: int, device: int = 0, halo_rows: int = HALO_ROWS, flags: Optional[int] = None,
                 group: Optional[dist.ProcessGroup] = None, use_table: bool = True, mode: str = "peer", row_weights: Optional[Sequence[float]] = None, settings=None):
        from . import executor as ex
        self.ex = ex
        self.rank, self.world, self.height, self.width, self.halo, self.group = rank, world, height, width, halo_rows, group
        self.strips = strip_rows(height, world, row_weights, halo_rows if world > 1 else 0)
        self.rows = self.strips[rank]
        if world > 1 and min(b - a for a, b in self.strips) < halo_rows:
            raise ValueError(f"strips of {min(b - a for a, b in self.strips)} rows are shorter than the {halo_rows}-row halo")
        self.den = ex.CudaDenoiser(denoiser, width, height, device=device, flags=ex.FLAG_QUAD_INTRINSICS if flags is None else flags)
        self.denoiser, self._settings = denoiser, settings
        # per-texture aprons of THIS denoiser ( binding indices differ between the REBLUR variants ); use_table=False -> `halo_rows` everywhere
        self.table = derive_halo_table(ex.host_library(), denoiser, width, height, halo_rows, settings) if use_table else {}
        if settings is not None:
            self.den.set_denoiser_settings(settings)
        self.device = device
        self.bytes_sent = 0
        self.use_table = use_table
        self._plans = {}
        self.mode = mode if world > 1 else "single"
        self._attached = False
        self._shared = []
        self.on_pass: Optional[Callable[[int, str], None]] = None
        self._cb = ex.DISPATCH_CALLBACK(self._after_dispatch)

    # the executor calls this after it has enqueued dispatch `index` on the stream
    def _after_dispatch(self, user, index, name, textures, is_storage, n):
        # pools ping-pong between two bindings at most, so the send / recv list of a (pass, pointers) pair is built once
        pass_name = name.decode() if name else ""
        wanted = [(i, halo_rows_for(self.table, pass_name, i, self.halo) if self.use_table else self.halo) for i in range(n) if is_storage[i]]
        wanted = [(i, h) for i, h in wanted if h > 0]
        key = (name, tuple(textures[i].data for i, _ in wanted))
        plan = self._plans.get(key)
        if plan is None:
            planes = [self.ex._as_byte_tensor(textures[i].data, textures[i].pitchBytes * textures[i].height, self.device).view(textures[i].height, textures[i].pitchBytes)
                      for i, _ in wanted]
            plan = self._plans[key] = build_exchange(planes, self.strips, self.rank, self.height, [h for _, h in wanted], self.group)
        ops, sent = plan
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        self.bytes_sent += sent
        if self.on_pass:
            self.on_pass(index, name.decode() if name else "")

    # ---- peer mode -----------------------------------------------------------------------------------------------
    def shared_texture(self, rtype: int, fmt: int) -> torch.Tensor:
        """Allocate a user texture the denoiser WRITES (OUT_*) inside the context, so that the neighbouring strips can map it, bind
        it to `rtype` and return a tensor view [H, W(, C)] of it (rows are 256-byte aligned: the view may be strided)."""
        from . import nrd_api as api
        ex = self.ex
        tex = ex.CuTexture()
        ex._check(ex.load().nrdcuAllocSharedTexture(self.den.ctx, int(fmt), self.width, self.height, C.byref(tex)), "nrdcuAllocSharedTexture")
        ex._check(ex.load().nrdcuSetResource(self.den.ctx, int(rtype), C.byref(tex)), "nrdcuSetResource")
        dtype, ch = ex.FORMAT_STORAGE[api.Format(fmt)]
        raw = ex._as_byte_tensor(tex.data, tex.pitchBytes * tex.height, self.device).view(dtype)
        pitch_elems = tex.pitchBytes // raw.element_size()
        view = raw.as_strided((self.height, self.width, ch), (pitch_elems, ch, 1)) if ch > 1 else raw.as_strided((self.height, self.width), (pitch_elems, 1))
        self._shared.append(view)
        return view

    def attach_peers(self):
        """Trade IPC handles with the neighbouring ranks and map their textures (collective: call on every rank, once, after all
        shared textures exist and before the first denoise)."""
        if self.mode != "peer" or self._attached:
            return
        ex, L = self.ex, self.ex.load()
        if self.use_table:
            for (pass_name, binding), rows in self.table.items():
                ex._check(L.nrdcuTileSetHalo(self.den.ctx, pass_name.encode(), binding, min(rows, self.halo)), "nrdcuTileSetHalo")
        ex._check(L.nrdcuTileSetHalo(self.den.ctx, None, 0, self.halo), "nrdcuTileSetHalo")
        size = L.nrdcuTileExportSize(self.den.ctx)
        blob = C.create_string_buffer(size)
        ex._check(L.nrdcuTileExport(self.den.ctx, blob, size), "nrdcuTileExport")
        blobs = [None] * self.world
        dist.all_gather_object(blobs, bytes(blob.raw), group=self.group)
        above = C.create_string_buffer(blobs[self.rank - 1], size) if self.rank > 0 else None
        below = C.create_string_buffer(blobs[self.rank + 1], size) if self.rank + 1 < self.world else None
        ex._check(L.nrdcuTileAttach(self.den.ctx, above, below, size), "nrdcuTileAttach")
        dist.barrier(group=self.group)   # nobody pushes before everybody has mapped
        self._attached = True

    def status(self):
        """(bytes pushed to the neighbours so far, error code of the flag wait: 0 = fine)."""
        b, e = C.c_uint64(), C.c_uint32()
        self.ex._check(self.ex.load().nrdcuTileGetStatus(self.den.ctx, C.byref(b), C.byref(e)), "nrdcuTileGetStatus")
        return int(b.value), int(e.value)

    def motion_bound(self):
        """(rows of vertical motion per frame the apron covers, worst overshoot any history fetch of this strip has had so far: 0 = fine).
        The executor checks every strip frame on the device ( nrdcuTileGetMotionBound ); reading the pair costs nothing."""
        bound, worst = C.c_uint32(), C.c_uint32()
        self.ex._check(self.ex.load().nrdcuTileGetMotionBound(self.den.ctx, C.byref(bound), C.byref(worst)), "nrdcuTileGetMotionBound")
        return int(bound.value), int(worst.value)

    def set_denoiser_settings(self, settings):
        """Blur radii / history-fix stride change how far passes reach: the apron table follows the settings ( before attach_peers in peer mode )."""
        if self.use_table and self.world > 1:
            if self._attached:
                raise RuntimeError("TiledDenoiser.set_denoiser_settings after attach_peers: pass `settings=` to the constructor instead")
            self.table = derive_halo_table(self.ex.host_library(), self.denoiser, self.width, self.height, self.halo, settings)
            self._plans.clear()
        self.den.set_denoiser_settings(settings)

    def __getattr__(self, item):  # set_user_texture, set_common_settings, set_denoiser_settings, pool_texture, profiling ...
        return getattr(self.den, item)

    def denoise(self, stream: Optional[torch.cuda.Stream] = None):
        s = stream.cuda_stream if stream is not None else torch.cuda.current_stream().cuda_stream
        L = self.ex.load()
        if self.mode == "peer" and not self._attached:
            self.attach_peers()
        cb = self._cb if self.mode == "nccl" else self.ex.DISPATCH_CALLBACK()
        self.ex._check(L.nrdcuDenoiseRows(self.den.ctx, self.den._ids, 1, C.c_void_p(s), self.rows[0], self.rows[1], cb, None), "nrdcuDenoiseRows")

    def close(self):
        self.den.close()
