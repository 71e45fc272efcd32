"""Comparison helpers shared by the parity tests. Tolerances are OURS (the reference states none):
fp16/fp32 planes: |a-b| <= atol + rtol*|b| on >= 99.9 % of texels and PSNR >= 60 dB for single-pass parity;
UNORM planes: exact or +-1 LSB; packed UINT planes: identical on >= 99.9 % of texels."""
import math

import torch

from nrd_sample_b200 import nrd_api as api


def decode(t: torch.Tensor, fmt: int, layout: str = "reblur") -> torch.Tensor:
    """Storage tensor -> float tensor (H, W, C) of decoded channel values (integers stay integers-as-float).
    `layout` selects the meaning of R32_UINT planes: "reblur" = REBLUR data2, "sigma" = asuint(viewZ) & ~7 | history length."""
    t = t.detach().cpu()
    f = api.Format(fmt)
    if f == api.Format.R32_UINT and layout == "sigma":
        v = t.to(torch.int64) & 0xFFFFFFFF
        z = ((v & 0xFFFFFFF8) - ((v & 0x80000000) << 1)).to(torch.int32).view(torch.float32)   # two's complement back to int32 bits
        return torch.stack([z, (v & 7).float()], -1)
    if f == api.Format.R10_G10_B10_A2_UNORM:
        v = t.to(torch.int64) & 0xFFFFFFFF
        return torch.stack([(v & 1023), (v >> 10) & 1023, (v >> 20) & 1023, (v >> 30) & 3], -1).float()
    if f == api.Format.R16_UNORM:
        return ((t.to(torch.int64) & 0xFFFF).float() / 65535.0).unsqueeze(-1)
    if f == api.Format.RGBA16_SNORM:
        return (t.float() / 32767.0).clamp_min(-1.0)
    if f == api.Format.R16_UINT:
        v = t.to(torch.int64) & 0xFFFF
        return torch.stack([v & 63, (v >> 6) & 63, (v >> 12) & 15], -1).float()
    if f == api.Format.R32_UINT:
        v = t.to(torch.int64) & 0xFFFFFFFF
        curv = ((v >> 16) & 0xFFFF).to(torch.int32).to(torch.int16).view(torch.float16).float()   # REBLUR data2: fp16 curvature in the top half
        return torch.stack([(v & 0xFF).float(), ((v >> 8) & 127).float(), ((v >> 15) & 1).float(), curv], -1)
    x = t.float()
    return x if x.dim() == 3 else x.unsqueeze(-1)


def compare(a: torch.Tensor, b: torch.Tensor, fmt: int, atol=1e-3, rtol=2 ** -9, layout: str = "reblur"):
    """a = under test, b = oracle. Returns dict(frac_bad, max_abs, psnr, n)."""
    f = api.Format(fmt)
    da, db = decode(a, fmt, layout), decode(b, fmt, layout)
    assert da.shape == db.shape, (da.shape, db.shape)
    both_nan = torch.isnan(da) & torch.isnan(db)
    da = torch.where(both_nan, torch.zeros_like(da), da)
    db = torch.where(both_nan, torch.zeros_like(db), db)
    diff = (da - db).abs()
    if f in (api.Format.R8_UNORM, api.Format.RG8_UNORM, api.Format.RGBA8_UNORM, api.Format.R10_G10_B10_A2_UNORM):
        bad = diff > 1.0
    elif f == api.Format.R16_UNORM:
        bad = diff > 16.5 / 65535.0       # 12 bits of a 16-bit UNORM: the weights upstream are fast-math fp32 against IEEE fp32, amplified by the spatial passes
    elif f == api.Format.RGBA16_SNORM:
        bad = diff > 8.5 / 32767.0
    elif f in (api.Format.R16_UINT, api.Format.R8_UINT):
        bad = diff > 0.0
    elif f == api.Format.R32_UINT and layout == "sigma":
        bad = diff > 0.0                                                     # viewZ bits and 3-bit history length: identical
    elif f == api.Format.R32_UINT:
        bad = diff > 0.0
        bad[..., 1] = diff[..., 1] > 1.0                                     # 7-bit virtual history amount: +-1 LSB
        bad[..., 3] = ~(diff[..., 3] <= 1e-4 + 2 ** -7 * db[..., 3].abs())   # curvature (fp16)
    else:
        bad = ~(diff <= atol + rtol * db.abs())
    bad_px = bad.any(-1)
    finite = torch.isfinite(diff)
    mse = (diff[finite] ** 2).mean().item() if finite.any() else 0.0
    peak = max(db[torch.isfinite(db)].abs().max().item() if torch.isfinite(db).any() else 1.0, 1e-6)
    psnr = 10.0 * math.log10(peak * peak / mse) if mse > 0 else float("inf")
    return dict(frac_bad=bad_px.float().mean().item(), max_abs=diff[finite].max().item() if finite.any() else 0.0, psnr=psnr, n=bad_px.numel())
