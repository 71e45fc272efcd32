"""Pins the oracle's restatement of the MathLib helpers (oracle/nrd_shared.h) against the REFERENCE's own MathLib:
live against oracle/_ref/libml_ref.so (ml.hlsli compiled as C++ from /root/reference) when it exists, and always against
tests/golden/ml_vectors.json, which tests/golden/make_ml_golden.py recorded from that library.
Integer/bit functions must be bit-exact; float functions must agree to <= 2 ulp (same formulas, same libm)."""
import ctypes as C
import json
import os
import struct

import numpy as np
import pytest

from oracle import runner

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ml_vectors.json")
F, U, PF = C.c_float, C.c_uint32, C.POINTER(C.c_float)

# NOT pinned against the C++ build of MathLib: Packing::RgbaToUint / UintToRgba. Compiled as C++ their uint4 conversions
# and per-lane shifts do not follow HLSL semantics (UintToRgba(0xD03C, 6, 6, 4, 0) returns all zeros there), and only the
# shaders ever call them; the oracle follows the HLSL definition and is checked by round-trip properties below.
# name -> (argument kinds, result kind); 'f' float, 'u' uint32, 'fN' float array in, 'oN' float array out
SPEC = {
    "LinearStep": ("fff", "f"), "SmoothStep01": ("f", "f"), "SmoothStep": ("fff", "f"), "Pow01": ("ff", "f"), "Sqrt01": ("f", "f"),
    "AcosApproxPositive": ("f", "f"), "PositiveRcp": ("f", "f"), "Rsqrt": ("f", "f"), "Sign": ("f", "f"),
    "GetRotator": ("f", "o4"), "CombineRotators": (("f4", "f4"), "o4"), "ScaleRotator": (("f4", "f", "f"), "o4"), "RotateVector2": (("f4", "f", "f"), "o2"),
    "ReconstructViewPosition": (("f", "f", "f4", "f", "f"), "o3"), "GetScreenUv": (("f16", "f3"), "o2"), "ColorClamp": ("fff", "f"),
    "GetModifiedRoughnessFromNormalVariance": (("f", "f3"), "f"),
    "GetBilinearFilter": ("ffff", "o4"), "GetBilinearCustomWeights": (("f", "f", "f", "f", "f4"), "o4"), "ApplyBilinearCustomWeights": (("f4", "f4"), "f"),
    "GetCatmullRomOrigin": ("ffff", "o2"), "Hash": ("u", "u"), "HashCombine": ("uu", "u"), "Zorder": ("uu", "u"), "CheckerBoard": ("uuu", "u"),
    "Weyl1D": (("f", "u"), "f"), "RngHash": ("uuu", "o4"), "GetSpecularLobeTanHalfAngle": ("ff", "f"), "GetSpecularDominantFactorG2": ("ff", "f"),
}


def _kinds(spec):
    return list(spec) if isinstance(spec, str) else list(spec)


def make_inputs(name, rng, n=64):
    kinds = _kinds(SPEC[name][0])
    rows = []
    for _ in range(n):
        row = []
        for k in kinds:
            if k == "f":
                row.append(float(np.float32(rng.uniform(-0.25, 1.5))))
            elif k == "u":
                row.append(int(rng.integers(0, 2 ** 32)))
            else:
                cnt = int(k[1:])
                row.append([float(np.float32(x)) for x in rng.uniform(-1.0, 1.0, cnt)])
        if name == "ColorClamp":
            row[1] = abs(row[1])  # sigma >= 0 in every call site; clamp(lo > hi) is implementation-defined
        if name == "ApplyBilinearCustomWeights":
            row[1] = [abs(x) for x in row[1]]  # weights are products of [0,1] factors; no catastrophic cancellation in the sum
        rows.append(row)
    return rows


def call(lib, prefix, name, row):
    kinds, res = _kinds(SPEC[name][0]), SPEC[name][1]
    fn = getattr(lib, prefix + name)
    args, argtypes = [], []
    for k, v in zip(kinds, row):
        if k == "f":
            args.append(F(v)); argtypes.append(F)
        elif k == "u":
            args.append(U(v)); argtypes.append(U)
        else:
            arr = (F * int(k[1:]))(*v)
            args.append(arr); argtypes.append(PF)
    out = None
    if res.startswith("o"):
        out = (F * int(res[1:]))()
        args.append(out); argtypes.append(PF)
        fn.restype = None
    else:
        fn.restype = F if res == "f" else U
    fn.argtypes = argtypes
    r = fn(*args)
    if out is not None:
        return [struct.unpack("<I", struct.pack("<f", x))[0] for x in out]
    return [struct.unpack("<I", struct.pack("<f", r))[0]] if res == "f" else [int(r)]


def ulp_diff(a, b):
    def key(u):
        return u if u < 0x80000000 else 0x80000000 - u
    return abs(key(a) - key(b))


def check(name, got, want):
    exact = SPEC[name][1] == "u" or name in ("RngHash", "Weyl1D", "GetBilinearFilter", "GetCatmullRomOrigin", "Sign", "ColorClamp")
    for g, w in zip(got, want):
        if exact:
            assert g == w, f"{name}: {g:#x} != {w:#x}"
        else:
            assert ulp_diff(g, w) <= 2, f"{name}: {g:#x} vs {w:#x} ({ulp_diff(g, w)} ulp)"


@pytest.mark.parametrize("name", sorted(SPEC))
def test_against_golden_vectors(name):
    golden = json.load(open(GOLDEN))
    lib = runner.lib()
    for row, want in golden[name]:
        check(name, call(lib, "orc_", name, row), want)


@pytest.mark.parametrize("name", sorted(SPEC))
def test_against_reference_mathlib(name):
    path = os.path.join(os.path.dirname(runner.REF_LIB_PATH), "libml_ref.so")
    runner.build()
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libml_ref.so not built (reference tree not mounted)")
    ref, lib = C.CDLL(path), runner.lib()
    rng = np.random.default_rng(1234)
    for row in make_inputs(name, rng, 200):
        check(name, call(lib, "orc_", name, row), call(ref, "ml_", name, row))


def test_half_conversion_matches_numpy():
    lib = runner.lib()
    lib.nrd_oracle_f32tof16.argtypes = [F]
    lib.nrd_oracle_f32tof16.restype = C.c_uint16
    lib.nrd_oracle_f16tof32.argtypes = [C.c_uint16]
    lib.nrd_oracle_f16tof32.restype = F
    rng = np.random.default_rng(7)
    vals = np.concatenate([rng.normal(0, 1, 2000), rng.normal(0, 1e-5, 2000), rng.normal(0, 3e4, 2000), [0.0, -0.0, 65504.0, 65519.9, 65520.0, 1e9, 6e-8, 2.98e-8, 3e-8, np.inf]]).astype(np.float32)
    for v in vals:
        assert lib.nrd_oracle_f32tof16(float(v)) == int(np.float16(v).view(np.uint16)), v
    for h in list(range(0, 0x7C00, 37)) + [0x8000, 0xFBFF, 0x0001, 0x03FF, 0x0400]:
        assert lib.nrd_oracle_f16tof32(h) == float(np.uint16(h).view(np.float16))


def test_internal_data_packing_round_trip():
    """PackInternalData / UnpackInternalData bit layout (REBLUR_Common.hlsli:13-38, REBLUR_Config.hlsli:62-65):
    6 bits diff frames | 6 bits spec frames | 4 bits material, LSB first."""
    lib = runner.lib()
    lib.orc_RgbaToUint664.argtypes = [PF]
    lib.orc_RgbaToUint664.restype = U
    lib.orc_UintToRgba664.argtypes = [U, PF]
    lib.orc_UintToRgba664.restype = None
    for d in (0, 1, 17, 62, 63):
        for s in (0, 5, 63):
            for m in (0, 1, 7, 15):
                p = lib.orc_RgbaToUint664((F * 4)(d / 63.0, s / 63.0, m / 15.0, m / 15.0))
                assert p == d | (s << 6) | (m << 12)
                out = (F * 4)()
                lib.orc_UintToRgba664(p, out)
                assert round(out[0] * 63) == d and round(out[1] * 63) == s and round(out[2] * 15) == m and out[3] == 0.0
