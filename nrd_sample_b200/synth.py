"""Seeded synthetic NRD inputs: the reference ships no pixel data, so tests and bench.py
feed the denoisers an analytic "Bistro-shaped" scene — ground plane, back wall, boxes, spheres, ~25 % sky —
rendered to exactly the textures NRDSample's path tracer would hand to NRD
(Shaders/TraceOpaque.cs.hlsl:614-657,738-757: IN_VIEWZ R32F, IN_NORMAL_ROUGHNESS R10G10B10A2 via
NRD_FrontEnd_PackNormalAndRoughness, IN_MV RGBA16F 2.5D, IN_*_RADIANCE_HITDIST RGBA16F via
REBLUR_FrontEnd_PackRadianceAndNormHitDist), plus the matching nrd::CommonSettings camera matrices.

Everything is torch tensor math so it runs on the CPU here and on the GPU for bench-sized frames; the
generator is test/bench plumbing, not part of the denoiser.
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass
from typing import Dict, Tuple

import torch

from . import nrd_api as api

SEED_BASE = 0x4E5244
SKY_VIEWZ = 1.0e6           # > CommonSettings::denoisingRange (5e5): exercises the 16x16 sky-tile early-out
HIT_DIST_PARAMS = (3.0, 0.1, 20.0)


@dataclass
class Camera:
    view_to_clip: Tuple[float, ...]      # column-major 4x4
    world_to_view: Tuple[float, ...]     # column-major 4x4
    position: Tuple[float, float, float]
    rotation: torch.Tensor               # 3x3 world->view (float64)
    tan_half_fov_y: float
    aspect: float


def make_camera(frame: float, width: int, height: int, period: int = 0) -> Camera:
    """D3D-style LH perspective, fovY 60 deg, reversed-Z infinite far. The camera drifts 0.02 units/frame along +x
    and yaws 0.1 deg/frame; with `period` > 0 the path is a closed loop so a short ring of frames can be replayed."""
    fov = math.radians(60.0)
    t = 1.0 / math.tan(fov / 2.0)
    a = width / height
    near = 0.1
    v2c = (t / a, 0, 0, 0, 0, t, 0, 0, 0, 0, 0, 1, 0, 0, near, 0)
    if period:
        ph = 2.0 * math.pi * (frame % period) / period
        px, yaw = 0.02 * period / (2 * math.pi) * math.sin(ph), math.radians(0.1 * period / (2 * math.pi)) * math.cos(ph)
    else:
        px, yaw = 0.02 * frame, math.radians(0.1 * frame)
    c, s = math.cos(yaw), math.sin(yaw)
    R = torch.tensor([[c, 0.0, -s], [0.0, 1.0, 0.0], [s, 0.0, c]], dtype=torch.float64)
    pos = torch.tensor([px, 1.5, -3.0], dtype=torch.float64)
    M = torch.eye(4, dtype=torch.float64)
    M[:3, :3] = R
    M[:3, 3] = -(R @ pos)
    w2v = tuple(float(x) for x in M.t().reshape(-1))
    return Camera(v2c, w2v, (float(pos[0]), float(pos[1]), float(pos[2])), R, math.tan(fov / 2.0), a)


def common_settings(frame_index: int, width: int, height: int, period: int = 0, **overrides) -> api.CommonSettings:
    cur = make_camera(frame_index, width, height, period)
    prev = make_camera(frame_index - 1 if frame_index > 0 or period else 0, width, height, period)
    cs = api.CommonSettings()
    cs.viewToClipMatrix = (C.c_float * 16)(*cur.view_to_clip)
    cs.viewToClipMatrixPrev = (C.c_float * 16)(*prev.view_to_clip)
    cs.worldToViewMatrix = (C.c_float * 16)(*cur.world_to_view)
    cs.worldToViewMatrixPrev = (C.c_float * 16)(*prev.world_to_view)
    cs.resourceSize = (C.c_uint16 * 2)(width, height)
    cs.resourceSizePrev = (C.c_uint16 * 2)(width, height)
    cs.rectSize = (C.c_uint16 * 2)(width, height)
    cs.rectSizePrev = (C.c_uint16 * 2)(width, height)
    cs.motionVectorScale = (C.c_float * 3)(1.0 / width, 1.0 / height, 1.0)
    cs.frameIndex = frame_index
    cs.timeDeltaBetweenFrames = 16.6   # removes the wall-clock dependence of gFramerateScale (InstanceImpl.cpp:439-443)
    for k, v in overrides.items():
        setattr(cs, k, v)
    return cs


# ------------------------------------------------------------------------------------------------
# Scene
# ------------------------------------------------------------------------------------------------
def _scene(shadow_casters: bool = False):
    """(kind, params, roughness, materialID). Fixed layout, units ~ metres. `shadow_casters` adds off-screen occluders
    (a floating slab and two beams) that only shadow rays see, so a good part of the ground lies in wide penumbra."""
    objs = [("plane_y", (0.0,), 0.5, 0), ("plane_z", (20.0, 7.5), 1.0, 0)]
    rough = (0.05, 0.2, 0.5, 1.0)
    for i in range(12):
        cx = -11.0 + 2.0 * i + 0.3 * math.sin(3.1 * i)
        cz = 4.0 + 1.3 * ((i * 7) % 11)
        h = 0.6 + 0.25 * ((i * 5) % 7)
        objs.append(("box", (cx - 0.6, 0.0, cz - 0.6, cx + 0.6, h, cz + 0.6), rough[i % 4], i % 2))
    for i in range(6):
        cx = -6.0 + 2.4 * i
        cz = 3.0 + 1.1 * ((i * 3) % 5)
        r = 0.45 + 0.1 * (i % 3)
        objs.append(("sphere", (cx, r, cz, r), rough[(i + 1) % 4], (i + 1) % 2))
    if shadow_casters:
        objs.append(("box", (-7.0, 3.0, 2.0, -1.0, 3.2, 9.0), 1.0, 0))
        objs.append(("box", (1.0, 5.0, 1.0, 1.4, 5.4, 16.0), 1.0, 0))
        objs.append(("box", (4.0, 2.0, 0.0, 9.0, 2.2, 0.6), 1.0, 0))
    return objs


def _intersect(o: torch.Tensor, d: torch.Tensor, shadow_casters: bool = False):
    """Nearest hit of the rays o + t d (o, d: H x W x 3) with the scene: (t or inf, normal, roughness, materialID)."""
    height, width = d.shape[0], d.shape[1]
    device, dtype = d.device, d.dtype
    inf = torch.full((height, width), float("inf"), device=device, dtype=dtype)
    tbest = inf.clone()
    nbest = torch.zeros(height, width, 3, device=device, dtype=dtype)
    rough = torch.ones(height, width, device=device, dtype=dtype)
    mat = torch.zeros(height, width, device=device, dtype=dtype)
    eps = 1e-9

    for kind, p, r, m in _scene(shadow_casters):
        if kind == "plane_y":
            t = (p[0] - o[..., 1]) / (d[..., 1] - eps)
            n = torch.tensor([0.0, 1.0, 0.0], device=device, dtype=dtype).expand(height, width, 3)
            ok = (t > 0) & (d[..., 1] < 0)
        elif kind == "plane_z":
            t = (p[0] - o[..., 2]) / (d[..., 2] + eps)
            hit_y = o[..., 1] + t * d[..., 1]
            n = torch.tensor([0.0, 0.0, -1.0], device=device, dtype=dtype).expand(height, width, 3)
            ok = (t > 0) & (hit_y < p[1])
        elif kind == "box":
            lo = torch.tensor(p[:3], device=device, dtype=dtype)
            hi = torch.tensor(p[3:], device=device, dtype=dtype)
            inv = 1.0 / torch.where(d.abs() < 1e-9, torch.full_like(d, 1e-9), d)
            t0 = (lo - o) * inv
            t1 = (hi - o) * inv
            tmin = torch.minimum(t0, t1)
            tmax = torch.maximum(t0, t1)
            tn, axis = tmin.max(-1)
            tf = tmax.min(-1).values
            ok = (tn < tf) & (tn > 0)
            t = tn
            n = torch.zeros(height, width, 3, device=device, dtype=dtype)
            sgn = -torch.sign(torch.gather(d, -1, axis.unsqueeze(-1)))
            n.scatter_(-1, axis.unsqueeze(-1), sgn)
        else:  # sphere
            c = torch.tensor(p[:3], device=device, dtype=dtype)
            oc = o - c
            dd = (d * d).sum(-1)
            b = (d * oc).sum(-1)
            cc = (oc * oc).sum(-1) - p[3] * p[3]
            disc = b * b - dd * cc
            ok = disc > 0
            t = (-b - torch.sqrt(disc.clamp_min(0))) / dd
            ok = ok & (t > 0)
            hitp = o + t.unsqueeze(-1) * d
            n = (hitp - c) / p[3]
        closer = ok & (t < tbest)
        tbest = torch.where(closer, t, tbest)
        nbest = torch.where(closer.unsqueeze(-1), n, nbest)
        rough = torch.where(closer, torch.full_like(rough, r), rough)
        mat = torch.where(closer, torch.full_like(mat, float(m)), mat)

    return tbest, nbest, rough, mat


def _raycast(cam: Camera, width: int, height: int, device, dtype=torch.float32):
    ys, xs = torch.meshgrid(torch.arange(height, device=device, dtype=dtype), torch.arange(width, device=device, dtype=dtype), indexing="ij")
    u = (xs + 0.5) / width
    v = (ys + 0.5) / height
    # view-space ray through the pixel centre, z = 1 (LH, +y up, uv.y down)
    dvx = (u * 2.0 - 1.0) * (cam.tan_half_fov_y * cam.aspect)
    dvy = (1.0 - v * 2.0) * cam.tan_half_fov_y
    Rt = cam.rotation.t().to(device=device, dtype=dtype)          # view->world
    dv = torch.stack([dvx, dvy, torch.ones_like(dvx)], -1)
    d = dv @ Rt.t()
    o = torch.tensor(cam.position, device=device, dtype=dtype)

    tbest, nbest, rough, mat = _intersect(o.expand(height, width, 3), d)

    hit = torch.isfinite(tbest)
    tsafe = torch.where(hit, tbest, torch.zeros_like(tbest))
    X = o + tsafe.unsqueeze(-1) * d
    viewz = torch.where(hit, tsafe, torch.full_like(tbest, SKY_VIEWZ))   # ray has view-space z = 1 => viewZ = t
    return dict(hit=hit, viewz=viewz, X=X, N=nbest, roughness=rough, material=mat, u=u, v=v, V=-d / d.norm(dim=-1, keepdim=True))


# ------------------------------------------------------------------------------------------------
# Packing (mirrors the front-end helpers of NRD.hlsli)
# ------------------------------------------------------------------------------------------------
def pack_normal_roughness(N: torch.Tensor, roughness: torch.Tensor, material: torch.Tensor) -> torch.Tensor:
    """NRD_FrontEnd_PackNormalAndRoughness (NRD.hlsli:696-722, encoding 2) + R10G10B10A2_UNORM quantisation -> int32 (H, W)."""
    n = N / (N.abs().sum(-1, keepdim=True) + 1e-20)
    ry = n[..., 1] * 0.5 + 0.5
    rx = n[..., 0] * 0.5 + ry
    ry = ry - n[..., 0] * 0.5
    rgh = roughness.clamp_min(1.5 / 512.0)
    s = torch.where(n[..., 2] < 0, -rgh, rgh)
    rz = s * 0.5 + 0.5
    a = (material / 3.0).clamp(0, 1)

    def q(x, m):
        return torch.floor(x.clamp(0, 1) * m + 0.5).to(torch.int64)

    word = q(rx, 1023.0) | (q(ry, 1023.0) << 10) | (q(rz, 1023.0) << 20) | (q(a, 3.0) << 30)
    word = torch.where(word >= 2 ** 31, word - 2 ** 32, word)
    return word.to(torch.int32)


def spec_magic_curve(roughness: torch.Tensor, power: float) -> torch.Tensor:
    return (1.0 - torch.exp2(-200.0 * roughness * roughness)) * roughness.clamp(0, 1).pow(power)


def hit_dist_normalization(viewz: torch.Tensor, roughness: torch.Tensor, params=HIT_DIST_PARAMS) -> torch.Tensor:
    """_REBLUR_GetHitDistanceNormalization (NRD.hlsli:568-573)."""
    smc = spec_magic_curve(roughness, 0.5)
    return (params[0] + viewz.abs() * params[1]) * (params[2] + (1.0 - params[2]) * smc)


def pack_radiance_hitdist(rgb: torch.Tensor, norm_hit_dist: torch.Tensor) -> torch.Tensor:
    """REBLUR_FrontEnd_PackRadianceAndNormHitDist(sanitize=true) (NRD.hlsli:815-826) -> float16 (H, W, 4)."""
    rgb = torch.nan_to_num(rgb, nan=0.0, posinf=65504.0).clamp(0, 65504.0)
    y = rgb[..., 0] * 0.25 + rgb[..., 1] * 0.5 + rgb[..., 2] * 0.25
    co = rgb[..., 0] * 0.5 - rgb[..., 2] * 0.5
    cg = -rgb[..., 0] * 0.25 + rgb[..., 1] * 0.5 - rgb[..., 2] * 0.25
    return torch.stack([y, co, cg, norm_hit_dist.clamp(0, 1)], -1).to(torch.float16)


def unpack_radiance(tex: torch.Tensor) -> torch.Tensor:
    """REBLUR_BackEnd_UnpackRadianceAndNormHitDist: YCoCg -> linear RGB (NRD.hlsli:418-428)."""
    t = tex.float()
    tt = t[..., 0] - t[..., 2]
    return torch.stack([tt + t[..., 1], t[..., 0] + t[..., 2], tt - t[..., 1]], -1).clamp_min(0)


# ------------------------------------------------------------------------------------------------
# Frame
# ------------------------------------------------------------------------------------------------
def checkerboard_pack(full: torch.Tensor, mode: int, frame_index: int) -> torch.Tensor:
    """Checkerboarded input as NRD expects it: a texture of the full resource size whose LEFT HALF holds the pixels the tracer produced this
    frame — texel (x >> 1, y) = full[y, x] for the pixels with Sequence::CheckerBoard( ( x, y ), frameIndex ) == mode (ml.hlsli:1620;
    CheckerboardMode::BLACK -> diffuse 0 / specular 1, WHITE -> diffuse 1 / specular 0, Reblur.cpp:301-313). The shaders address it both with
    integer loads (pos.x >> 1) and with uv.x * 0.5 (RELAX_PrePass.cs.hlsl:167), hence the full width. The width must be even."""
    h, w = full.shape[0], full.shape[1]
    assert w % 2 == 0
    ys = torch.arange(h, device=full.device)
    b = (mode ^ ys ^ frame_index) & 1                                      # x parity carrying data in row y
    xs = 2 * torch.arange(w // 2, device=full.device)[None, :] + b[:, None]
    out = torch.zeros_like(full)
    out[:, : w // 2] = full[ys[:, None], xs]
    return out.contiguous()


def reblur_frame(frame_index: int, width: int, height: int, device="cpu", period: int = 0, with_clean: bool = False, holes: bool = False,
                 checkerboard: int = 0, guides: bool = False, sh: bool = False) -> Dict[str, torch.Tensor]:
    """All user inputs of REBLUR_DIFFUSE_SPECULAR for one frame, in their API storage formats. `holes`: probabilistic lobe sampling as in
    NRDSample — every pixel traced only one lobe this frame (checkerboard flipping per frame), the other lobe has hit distance 0 and
    relies on ReblurSettings::hitDistanceReconstructionMode."""
    device = torch.device(device)
    cam = make_camera(frame_index, width, height, period)
    cam_prev = make_camera(frame_index - 1 if frame_index > 0 or period else 0, width, height, period)
    g = _raycast(cam, width, height, device)
    hit, viewz, X, N, rough, mat = g["hit"], g["viewz"], g["X"], g["N"], g["roughness"], g["material"]

    # 2.5D motion: where was this (static) world point on screen last frame, and how far along view z
    Rp = cam_prev.rotation.to(device=device, dtype=torch.float32)
    pp = torch.tensor(cam_prev.position, device=device, dtype=torch.float32)
    Xvp = (X - pp) @ Rp.t()
    zp = Xvp[..., 2].clamp_min(1e-6)
    up = (Xvp[..., 0] / zp / (cam_prev.tan_half_fov_y * cam_prev.aspect)) * 0.5 + 0.5
    vp = 0.5 - (Xvp[..., 1] / zp / cam_prev.tan_half_fov_y) * 0.5
    mvx = torch.where(hit, (up - g["u"]) * width, torch.zeros_like(up))
    mvy = torch.where(hit, (vp - g["v"]) * height, torch.zeros_like(vp))
    mvz = torch.where(hit, Xvp[..., 2] - viewz, torch.zeros_like(up))
    mv = torch.stack([mvx, mvy, mvz, torch.zeros_like(mvx)], -1).to(torch.float16)

    # clean signals: smooth lambert + a view-dependent lobe term, modulated by a smooth world-space pattern
    L = torch.tensor([0.4, 0.8, -0.45], device=device, dtype=torch.float32)
    L = L / L.norm()
    ndl = (N * L).sum(-1).clamp_min(0)
    checker = 0.75 + 0.25 * torch.sin(X[..., 0] * 0.9) * torch.cos(X[..., 2] * 0.7)   # smooth "lighting" variation (inputs are albedo-demodulated)
    tint_d = torch.tensor([1.0, 0.9, 0.75], device=device)
    tint_s = torch.tensor([0.8, 0.9, 1.0], device=device)
    diff_clean = ((ndl + 0.1) * checker).unsqueeze(-1) * tint_d
    V = g["V"]
    R = 2.0 * (N * V).sum(-1, keepdim=True) * N - V
    rdl = (R * L).sum(-1).clamp_min(0)
    shin = 2.0 / (rough * rough).clamp_min(1e-3)
    spec_clean = (rdl.pow(shin.clamp_max(64.0)) * 2.0 + 0.05).unsqueeze(-1) * tint_s

    gen = torch.Generator(device=device)
    gen.manual_seed(SEED_BASE + frame_index)

    def rnd():
        return torch.rand(height, width, device=device, generator=gen)

    def noisy(clean):
        e = -torch.log(1.0 - rnd() * 0.999999)                 # Exp(1): 1-spp-like multiplicative noise, mean 1
        fire = torch.where(rnd() < 0.002, torch.full_like(e, 50.0), torch.ones_like(e))
        return clean * (e * fire).unsqueeze(-1)

    diff_noisy = noisy(diff_clean)
    spec_noisy = noisy(spec_clean)
    hit_t_d = (0.1 + 9.9 * rnd()) * 2.0
    hit_t_s = (0.1 + 9.9 * rnd()) * (1.0 + rough)
    nhd_d = (hit_t_d / hit_dist_normalization(viewz, torch.ones_like(rough))).clamp(0, 1)
    nhd_s = (hit_t_s / hit_dist_normalization(viewz, rough)).clamp(0, 1)

    if holes:
        ys, xs = torch.meshgrid(torch.arange(height, device=device), torch.arange(width, device=device), indexing="ij")
        checker = ((xs + ys + frame_index) & 1).bool()
        nhd_d = torch.where(checker, nhd_d, torch.zeros_like(nhd_d))
        nhd_s = torch.where(checker, torch.zeros_like(nhd_s), nhd_s)

    zero4 = torch.zeros(height, width, 4, device=device, dtype=torch.float16)
    out = {
        "IN_VIEWZ": viewz.contiguous(),
        "IN_NORMAL_ROUGHNESS": pack_normal_roughness(N, rough, mat).contiguous(),
        "IN_MV": mv.contiguous(),
        "IN_DIFF_RADIANCE_HITDIST": torch.where(hit[..., None], pack_radiance_hitdist(diff_noisy, nhd_d), zero4).contiguous(),
        "IN_SPEC_RADIANCE_HITDIST": torch.where(hit[..., None], pack_radiance_hitdist(spec_noisy, nhd_s), zero4).contiguous(),
    }
    if guides:
        # Optional single-channel guides ( CommonSettings::isHistoryConfidenceAvailable / isDisocclusionThresholdMixAvailable ). NRD declares them
        # Texture2D<float> and samples the confidences by uv, so size and format are the application's: NRDSample binds an RGBA16F texture at
        # SHARC resolution ( Source/NRDSample.cpp:457-462, 2990 ). Here: diffuse confidence RGBA16F at half resolution, specular R8 at full
        # resolution, the threshold mix R16F — three different formats through the same kernel path.
        hw, hh = (width + 1) // 2, (height + 1) // 2
        yy, xx = torch.meshgrid(torch.arange(hh, device=device), torch.arange(hw, device=device), indexing="ij")
        conf_d = (0.55 + 0.45 * torch.sin(0.37 * xx + 0.11 * frame_index) * torch.cos(0.23 * yy)).clamp(0, 1)
        out["IN_DIFF_CONFIDENCE"] = torch.stack([conf_d, 1.0 - conf_d, torch.zeros_like(conf_d), torch.ones_like(conf_d)], -1).to(torch.float16).contiguous()
        conf_s = (rnd() * 1.3).clamp(0, 1)
        out["IN_SPEC_CONFIDENCE"] = (conf_s * 255.0 + 0.5).to(torch.uint8).contiguous()
        out["IN_DISOCCLUSION_THRESHOLD_MIX"] = (rough > 0.3).to(torch.float16).contiguous()
    if sh:
        # REBLUR_DIFFUSE_SPECULAR_SH ( NRD_MODE = SH ), REBLUR_FrontEnd_PackSh ( NRD.hlsli:831-849 ): SH0 = the same { Y, Co, Cg, normHitDist }, SH1 =
        # { direction * Y, 0 }. Directions as in `relax_frame` ( around N / the mirror direction ), from their own generator so that the
        # RADIANCE inputs of a frame do not depend on `sh`.
        gen_sh = torch.Generator(device=device)
        gen_sh.manual_seed(SEED_BASE + 0x5348 + frame_index)

        def around(axis, spread):
            d = axis + spread.unsqueeze(-1) * (torch.rand(height, width, 3, device=device, generator=gen_sh) * 2.0 - 1.0)
            return d / d.norm(dim=-1, keepdim=True).clamp_min(1e-6)

        for lobe, axis, spread in (("DIFF", N, torch.full_like(rough, 0.8)), ("SPEC", R, rough * 0.7 + 0.02)):
            sh0 = out.pop(f"IN_{lobe}_RADIANCE_HITDIST")
            out[f"IN_{lobe}_SH0"] = sh0
            sh1 = torch.cat([around(axis, spread) * sh0[..., :1].float(), torch.zeros(height, width, 1, device=device)], -1).to(torch.float16)
            out[f"IN_{lobe}_SH1"] = torch.where(hit[..., None], sh1, zero4).contiguous()
    if checkerboard:   # nrd::CheckerboardMode: 1 = BLACK, 2 = WHITE; SH1 travels with its SH0
        diff_mode, spec_mode = (0, 1) if checkerboard == 1 else (1, 0)
        for k in (("IN_DIFF_SH0", "IN_DIFF_SH1") if sh else ("IN_DIFF_RADIANCE_HITDIST",)):
            out[k] = checkerboard_pack(out[k], diff_mode, frame_index)
        for k in (("IN_SPEC_SH0", "IN_SPEC_SH1") if sh else ("IN_SPEC_RADIANCE_HITDIST",)):
            out[k] = checkerboard_pack(out[k], spec_mode, frame_index)
    if with_clean:
        out["_clean_diff"] = diff_clean
        out["_clean_spec"] = spec_clean
        out["_hit"] = hit
    return out


def occlusion_frame(frame_index: int, width: int, height: int, device="cpu", period: int = 0, lobes: str = "both", directional: bool = False, checkerboard: int = 0,
                    guides: bool = False, holes: bool = False, rgba16f: bool = False) -> Dict[str, torch.Tensor]:
    """User inputs of the REBLUR occlusion denoisers: the G-buffer and motion of `reblur_frame` with the normalized hit distance alone.
    REBLUR_*_OCCLUSION: IN_DIFF_HITDIST / IN_SPEC_HITDIST as R16_UNORM ( int16 storage ), or — `rgba16f`, what NRDSample binds ( Source/NRDSample.cpp:495-500 ) —
    as RGBA16F with the value in .x. REBLUR_DIFFUSE_DIRECTIONAL_OCCLUSION ( `directional` ): IN_DIFF_DIRECTION_HITDIST RGBA16F = { direction * hitDist, hitDist },
    the shape of REBLUR_FrontEnd_PackDirectionalOcclusion ( NRD.hlsli:853-864 ). Checkerboarded inputs are half width in the left half, as for `reblur_frame`."""
    base = reblur_frame(frame_index, width, height, device, period, holes=holes, guides=guides)
    out = {k: v for k, v in base.items() if "RADIANCE" not in k}
    nhd = {"DIFF": base["IN_DIFF_RADIANCE_HITDIST"][..., 3].float(), "SPEC": base["IN_SPEC_RADIANCE_HITDIST"][..., 3].float()}
    if directional:
        cam = make_camera(frame_index, width, height, period)
        n = _raycast(cam, width, height, torch.device(device))["N"]
        d = torch.cat([n * nhd["DIFF"].unsqueeze(-1), nhd["DIFF"].unsqueeze(-1)], -1).to(torch.float16)
        out["IN_DIFF_DIRECTION_HITDIST"] = (checkerboard_pack(d, 0 if checkerboard == 1 else 1, frame_index) if checkerboard else d).contiguous()
        return out
    diff_mode, spec_mode = (0, 1) if checkerboard == 1 else (1, 0)
    for lobe, mode in (("DIFF", diff_mode), ("SPEC", spec_mode)):
        if lobes != "both" and lobes.upper() != lobe[:4]:
            continue
        v = nhd[lobe].clamp(0, 1)
        if rgba16f:
            t = torch.stack([v, torch.zeros_like(v), torch.zeros_like(v), torch.ones_like(v)], -1).to(torch.float16)
        else:
            t = (v * 65535.0 + 0.5).to(torch.int32).clamp(0, 65535).to(torch.uint16).view(torch.int16)
        out[f"IN_{lobe}_HITDIST"] = (checkerboard_pack(t, mode, frame_index) if checkerboard else t).contiguous()
    if lobes == "diff":
        out.pop("IN_SPEC_CONFIDENCE", None)
    if lobes == "spec":
        out.pop("IN_DIFF_CONFIDENCE", None)
    return out


def reblur_frame_sh(frame_index: int, width: int, height: int, device="cpu", period: int = 0) -> Dict[str, torch.Tensor]:
    """The inputs of REBLUR_DIFFUSE_SPECULAR_SH ( `reblur_frame( sh = True )` ) under a name `bench.py --denoiser reblur_sh` can look up."""
    return reblur_frame(frame_index, width, height, device, period, sh=True)


def reference_frame(frame_index: int, width: int, height: int, device="cpu", period: int = 0) -> Dict[str, torch.Tensor]:
    """IN_SIGNAL of the REFERENCE denoiser: the noisy diffuse radiance of `reblur_frame` as an RGBA16F image (NRDSample feeds its composed frame)."""
    return {"IN_SIGNAL": reblur_frame(frame_index, width, height, device, period)["IN_DIFF_RADIANCE_HITDIST"]}


# ------------------------------------------------------------------------------------------------
# SIGMA_SHADOW inputs
# ------------------------------------------------------------------------------------------------
SIGMA_LIGHT_DIRECTION = (0.6, 0.5, 0.2)       # direction TO the light (not normalised)
SIGMA_TAN_ANGULAR_RADIUS = math.tan(math.radians(5.0))
FP16_MAX = 65504.0


def sigma_frame(frame_index: int, width: int, height: int, device="cpu", period: int = 0, with_clean: bool = False, translucency: bool = False) -> Dict[str, torch.Tensor]:
    """All user inputs of SIGMA_SHADOW for one frame: IN_VIEWZ, IN_NORMAL_ROUGHNESS, IN_MV as for REBLUR plus IN_PENUMBRA
    (R16F) = SIGMA_FrontEnd_PackPenumbra (NRD.hlsli:974-980) of a 1-spp shadow ray towards a disk light: 0 where the
    surface faces away from the light, FP16_MAX where the ray escapes, distance-to-occluder * tan(angular radius) / 2 otherwise."""
    device = torch.device(device)
    base = reblur_frame(frame_index, width, height, device, period)
    cam = make_camera(frame_index, width, height, period)
    g = _raycast(cam, width, height, device)
    hit, X, N = g["hit"], g["X"], g["N"]

    gen = torch.Generator(device=device)
    gen.manual_seed(SEED_BASE + 0x5147 + frame_index)
    L = torch.tensor(SIGMA_LIGHT_DIRECTION, device=device, dtype=torch.float32)
    L = L / L.norm()
    # orthonormal frame around L, uniform sample of the light disk
    up = torch.tensor([0.0, 0.0, 1.0], device=device)
    T = torch.linalg.cross(up, L)
    T = T / T.norm()
    B = torch.linalg.cross(L, T)
    o = X + N * 1e-3

    def shadow_ray():
        r = torch.sqrt(torch.rand(height, width, device=device, generator=gen)) * SIGMA_TAN_ANGULAR_RADIUS
        phi = torch.rand(height, width, device=device, generator=gen) * (2.0 * math.pi)
        d = L + (r * torch.cos(phi)).unsqueeze(-1) * T + (r * torch.sin(phi)).unsqueeze(-1) * B
        d = d / d.norm(dim=-1, keepdim=True)
        return _intersect(o, d, shadow_casters=True)[0]

    t = shadow_ray()
    ndl = (N * L).sum(-1)
    dist = torch.where(torch.isfinite(t), t, torch.full_like(t, FP16_MAX))
    penumbra = torch.where(dist >= FP16_MAX, torch.full_like(dist, FP16_MAX), (dist * SIGMA_TAN_ANGULAR_RADIUS * 0.5).clamp_max(32768.0))
    penumbra = torch.where(ndl <= 0.0, torch.zeros_like(penumbra), penumbra)
    penumbra = torch.where(hit, penumbra, torch.full_like(penumbra, FP16_MAX))
    out = {
        "IN_VIEWZ": base["IN_VIEWZ"],
        "IN_NORMAL_ROUGHNESS": base["IN_NORMAL_ROUGHNESS"],
        "IN_MV": base["IN_MV"],
        "IN_PENUMBRA": penumbra.to(torch.float16).contiguous(),
    }
    if translucency:
        # SIGMA_SHADOW_TRANSLUCENCY: IN_TRANSLUCENCY (RGBA8, as NRDSample creates it) = SIGMA_FrontEnd_PackTranslucency (NRD.hlsli:994-1001):
        # x = the shadow ray escaped, yzw = tint of the occluder it went through: 40 % of the blocked rays hit a "stained glass" occluder
        lit = dist >= FP16_MAX
        u = torch.rand(height, width, device=device, generator=gen)
        glass = (~lit) & (torch.rand(height, width, device=device, generator=gen) < 0.4)
        tint = torch.tensor([0.9, 0.45, 0.2], device=device) * (0.5 + 0.5 * u).unsqueeze(-1)
        rgb = torch.where(lit.unsqueeze(-1), torch.ones_like(tint), torch.where(glass.unsqueeze(-1), tint, torch.zeros_like(tint)))
        packed = torch.cat([lit.float().unsqueeze(-1), rgb.clamp(0.0, 1.0)], dim=-1)
        out["IN_TRANSLUCENCY"] = (packed * 255.0 + 0.5).to(torch.uint8).contiguous()
    if with_clean:   # converged visibility: mean over 48 more light samples
        vis = torch.zeros_like(ndl)
        for _ in range(48):
            vis += (~torch.isfinite(shadow_ray())).float()
        out["_clean_visibility"] = torch.where(ndl <= 0.0, torch.zeros_like(vis), vis / 48.0)
        out["_hit"] = hit
    return out


# ------------------------------------------------------------------------------------------------
# RELAX_DIFFUSE_SPECULAR_SH inputs
# ------------------------------------------------------------------------------------------------
def relax_frame(frame_index: int, width: int, height: int, device="cpu", period: int = 0, with_clean: bool = False, sh: bool = True, checkerboard: int = 0,
                guides: bool = False, holes: bool = False) -> Dict[str, torch.Tensor]:
    """All user inputs of RELAX_DIFFUSE_SPECULAR_SH for one frame (BASELINE.json config 2): the G-buffer and motion of `reblur_frame`,
    un-normalised radiance + hit distance in IN_*_SH0 and `direction * luminance` in IN_*_SH1, both RGBA16F, packed like
    RELAX_FrontEnd_PackSh (NRD.hlsli:925-941). Directions: cosine-weighted around N (diffuse), jittered mirror direction (specular)."""
    device = torch.device(device)
    cam = make_camera(frame_index, width, height, period)
    g = _raycast(cam, width, height, device)
    base = reblur_frame(frame_index, width, height, device, period, with_clean=True, guides=guides)
    hit, N, V, rough = g["hit"], g["N"], g["V"], g["roughness"]

    gen = torch.Generator(device=device)
    gen.manual_seed(SEED_BASE + 0x52454C + frame_index)

    def rnd():
        return torch.rand(height, width, device=device, generator=gen)

    def noisy(clean):
        # RELAX stores squared luminance (2nd moments) in fp16: radiance is capped at 200 the way a renderer caps fireflies before
        # handing them to RELAX, so that luminance^2 stays below 65504 (an Inf there would spread through the a-trous passes)
        e = -torch.log(1.0 - rnd() * 0.999999)
        fire = torch.where(rnd() < 0.002, torch.full_like(e, 50.0), torch.ones_like(e))
        return (clean * (e * fire).unsqueeze(-1)).clamp_max(200.0)

    def around(axis, spread):
        d = axis + spread.unsqueeze(-1) * (torch.stack([rnd(), rnd(), rnd()], -1) * 2.0 - 1.0)
        return d / d.norm(dim=-1, keepdim=True).clamp_min(1e-6)

    diff = noisy(base["_clean_diff"])
    spec = noisy(base["_clean_spec"])
    R = 2.0 * (N * V).sum(-1, keepdim=True) * N - V
    dir_d = around(N, torch.full_like(rough, 0.8))
    dir_s = around(R, rough * 0.7 + 0.02)
    hit_t_d = (0.1 + 9.9 * rnd()) * 2.0
    hit_t_s = (0.1 + 9.9 * rnd()) * (1.0 + rough)
    lum = torch.tensor([0.2126, 0.7152, 0.0722], device=device)
    zero4 = torch.zeros(height, width, 4, device=device, dtype=torch.float16)

    def sh0(rad, t):
        return torch.where(hit[..., None], torch.cat([rad.clamp(0, FP16_MAX), t.unsqueeze(-1)], -1).to(torch.float16), zero4).contiguous()

    def sh1(rad, d):
        return torch.where(hit[..., None], torch.cat([d * (rad * lum).sum(-1, keepdim=True), torch.zeros_like(d[..., :1])], -1).to(torch.float16), zero4).contiguous()

    out = {
        "IN_VIEWZ": base["IN_VIEWZ"],
        "IN_NORMAL_ROUGHNESS": base["IN_NORMAL_ROUGHNESS"],
        "IN_MV": base["IN_MV"],
        "IN_DIFF_SH0": sh0(diff, hit_t_d),
        "IN_DIFF_SH1": sh1(diff, dir_d),
        "IN_SPEC_SH0": sh0(spec, hit_t_s),
        "IN_SPEC_SH1": sh1(spec, dir_s),
    }
    if holes:   # probabilistic lobe sampling: each pixel traced one lobe, the other one has no hit distance ( HitDistanceReconstructionMode )
        ys, xs = torch.meshgrid(torch.arange(height, device=device), torch.arange(width, device=device), indexing="ij")
        checker = ((xs + ys + frame_index) & 1).bool()
        out["IN_DIFF_SH0"][..., 3] = torch.where(checker, out["IN_DIFF_SH0"][..., 3], torch.zeros_like(out["IN_DIFF_SH0"][..., 3]))
        out["IN_SPEC_SH0"][..., 3] = torch.where(checker, torch.zeros_like(out["IN_SPEC_SH0"][..., 3]), out["IN_SPEC_SH0"][..., 3])
    if guides:
        for k in ("IN_DIFF_CONFIDENCE", "IN_SPEC_CONFIDENCE", "IN_DISOCCLUSION_THRESHOLD_MIX"):
            out[k] = base[k]
    if checkerboard:   # nrd::CheckerboardMode: 1 = BLACK, 2 = WHITE; SH1 travels with its SH0
        diff_mode, spec_mode = (0, 1) if checkerboard == 1 else (1, 0)
        for k in ("IN_DIFF_SH0", "IN_DIFF_SH1"):
            out[k] = checkerboard_pack(out[k], diff_mode, frame_index)
        for k in ("IN_SPEC_SH0", "IN_SPEC_SH1"):
            out[k] = checkerboard_pack(out[k], spec_mode, frame_index)
    if not sh:   # RELAX_DIFFUSE_SPECULAR (NRD_MODE = RADIANCE): the same radiance + hit distance, no SH1 textures
        out["IN_DIFF_RADIANCE_HITDIST"], out["IN_SPEC_RADIANCE_HITDIST"] = out.pop("IN_DIFF_SH0"), out.pop("IN_SPEC_SH0")
        del out["IN_DIFF_SH1"], out["IN_SPEC_SH1"]
    if with_clean:
        out["_clean_diff"], out["_clean_spec"], out["_hit"] = base["_clean_diff"], base["_clean_spec"], hit
    return out
