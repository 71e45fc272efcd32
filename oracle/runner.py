"""TEST INFRASTRUCTURE — drives the CPU oracle (oracle/liboracle.so) through a whole nrd::Instance frame.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this module.
It plays the role NRDIntegration plays for the reference (External/NRD/Integration/NRDIntegration.hpp:522-890):
owns the permanent/transient pools, resolves each DispatchDesc's bindings to pool or user textures and replays the
dispatches in order — on host memory, through `nrd_oracle_dispatch`.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Callable, Dict, List, Optional, Sequence

import torch

from nrd_sample_b200 import nrd_api as api

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")
REF_LIB_PATH = os.path.join(HERE, "_ref", "libnrd_ref.so")
# the reference's own compute shaders compiled as C++ ( oracle/ref_build_shaders.py ): what pins the oracle's pixel math
REF_SHADERS_PATH = os.path.join(HERE, "_ref", "libnrd_refshaders.so")


class OracleTexture(C.Structure):
    _fields_ = [("data", C.c_void_p), ("width", C.c_uint32), ("height", C.c_uint32), ("pitchBytes", C.c_uint32), ("format", C.c_uint32)]


def build(force: bool = False) -> str:
    """Compile the oracle (and, when the reference tree is mounted, oracle/_ref). Returns the .so path."""
    srcs = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".cpp", ".h"))]
    srcs += [os.path.join(HERE, "..", "include", "nrd_frontend.cuh"), os.path.join(HERE, "..", "nrd_sample_b200", "csrc", "frontend_probe.inl")]   # frontend_probe.cpp
    stale = force or not os.path.exists(LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs)
    if stale:
        subprocess.check_call(["make", "-C", HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    if os.path.isdir("/root/reference/External/NRD/Source") and not os.path.exists(REF_LIB_PATH):
        subprocess.check_call([os.path.join(HERE, "ref_build.sh")], stdout=subprocess.DEVNULL)
    if os.path.isdir("/root/reference/External/NRD/Shaders"):
        shim = os.path.join(HERE, "ref_shim")
        deps = [os.path.join(shim, f) for f in os.listdir(shim)] + [os.path.join(HERE, "ref_build_shaders.py")]
        deps += [os.path.join(shim, "Shaders", f) for f in os.listdir(os.path.join(shim, "Shaders"))]
        if force or not os.path.exists(REF_SHADERS_PATH) or any(os.path.getmtime(d) > os.path.getmtime(REF_SHADERS_PATH) for d in deps):
            import sys
            subprocess.check_call([sys.executable, os.path.join(HERE, "ref_build_shaders.py")], stdout=subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB_PATH)
        _lib.nrd_oracle_dispatch.argtypes = [C.c_char_p, C.c_void_p, C.c_uint32, C.POINTER(OracleTexture), C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
        _lib.nrd_oracle_dispatch.restype = C.c_int
        _lib.nrd_oracle_set_threads.argtypes = [C.c_int]
        _lib.nrd_oracle_max_threads.restype = C.c_int
    return _lib


_ref_shaders = None


def ref_shaders() -> Optional[C.CDLL]:
    """libnrd_refshaders.so (prebuilt in this container from /root/reference; travels to the GPU box), or None when it was never built."""
    global _ref_shaders
    if _ref_shaders is None:
        build()
        if not os.path.exists(REF_SHADERS_PATH):
            return None
        _ref_shaders = C.CDLL(REF_SHADERS_PATH)
        _ref_shaders.nrd_refshader_dispatch.argtypes = [C.c_char_p, C.c_void_p, C.c_uint32, C.POINTER(OracleTexture), C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32]
        _ref_shaders.nrd_refshader_dispatch.restype = C.c_int
        _ref_shaders.nrd_refshader_name.restype = C.c_char_p
        _ref_shaders.nrd_refshader_name.argtypes = [C.c_int]
        _ref_shaders.nrd_refshader_set_threads.argtypes = [C.c_int]
        if hasattr(_ref_shaders, "nrd_refshader_probe"):
            _ref_shaders.nrd_refshader_probe.argtypes = [C.POINTER(C.c_uint64), C.c_int]
            _ref_shaders.nrd_refshader_probe.restype = None
    return _ref_shaders


def ref_mirror_probe(reset: bool = False):
    """(taps, taps that took the "mirrored" branch of REBLUR_Common_SpatialFilter.hlsli:198) counted by the reference shaders since the last reset."""
    out = (C.c_uint64 * 14)()
    ref_shaders().nrd_refshader_probe(out, 1 if reset else 0)
    return int(out[0]), int(out[1])


def ref_mirror_probe_detail():
    """{(pass, lobe): (taps, mirrored)} of the same counters ( see executor.mirror_probe_detail )."""
    out = (C.c_uint64 * 14)()
    ref_shaders().nrd_refshader_probe(out, 0)
    return {(p, l): (int(out[2 + 2 * (i * 2 + j)]), int(out[3 + 2 * (i * 2 + j)])) for i, p in enumerate(("Pre-pass", "Blur", "Post-blur")) for j, l in enumerate(("diff", "spec"))}


def ref_shader_names() -> List[str]:
    L = ref_shaders()
    return [L.nrd_refshader_name(i).decode() for i in range(L.nrd_refshader_count())] if L else []


# nrd::Format -> (torch dtype, channels)
FORMAT_STORAGE = {
    api.Format.R8_UNORM: (torch.uint8, 1),
    api.Format.R8_UINT: (torch.uint8, 1),
    api.Format.RG8_UNORM: (torch.uint8, 2),
    api.Format.RGBA8_UNORM: (torch.uint8, 4),
    api.Format.R16_UINT: (torch.int16, 1),
    api.Format.R16_SFLOAT: (torch.float16, 1),
    api.Format.R16_UNORM: (torch.int16, 1),
    api.Format.RGBA16_SNORM: (torch.int16, 4),
    api.Format.RGBA16_SFLOAT: (torch.float16, 4),
    api.Format.R32_UINT: (torch.int32, 1),
    api.Format.R32_SFLOAT: (torch.float32, 1),
    api.Format.R10_G10_B10_A2_UNORM: (torch.int32, 1),
    api.Format.RGBA32_SFLOAT: (torch.float32, 4),
}

# Formats of the user-provided resources as NRDSample creates them (Source/NRDSample.cpp:2913-3002)
USER_FORMATS = {
    api.ResourceType.IN_MV: api.Format.RGBA16_SFLOAT,
    api.ResourceType.IN_NORMAL_ROUGHNESS: api.Format.R10_G10_B10_A2_UNORM,
    api.ResourceType.IN_VIEWZ: api.Format.R32_SFLOAT,
    api.ResourceType.IN_DIFF_RADIANCE_HITDIST: api.Format.RGBA16_SFLOAT,
    api.ResourceType.IN_SPEC_RADIANCE_HITDIST: api.Format.RGBA16_SFLOAT,
    api.ResourceType.OUT_DIFF_RADIANCE_HITDIST: api.Format.RGBA16_SFLOAT,
    api.ResourceType.OUT_SPEC_RADIANCE_HITDIST: api.Format.RGBA16_SFLOAT,
    api.ResourceType.IN_DIFF_HITDIST: api.Format.R16_UNORM,            # "R8+" ( NRDDescs.h:77-80 ); NRDSample binds RGBA16F ( pass fmt= explicitly )
    api.ResourceType.IN_SPEC_HITDIST: api.Format.R16_UNORM,
    api.ResourceType.OUT_DIFF_HITDIST: api.Format.R16_UNORM,
    api.ResourceType.OUT_SPEC_HITDIST: api.Format.R16_UNORM,
    api.ResourceType.IN_DIFF_DIRECTION_HITDIST: api.Format.RGBA16_SFLOAT,
    api.ResourceType.OUT_DIFF_DIRECTION_HITDIST: api.Format.RGBA16_SFLOAT,
    api.ResourceType.IN_DIFF_SH0: api.Format.RGBA16_SFLOAT,
    api.ResourceType.IN_DIFF_SH1: api.Format.RGBA16_SFLOAT,
    api.ResourceType.IN_SPEC_SH0: api.Format.RGBA16_SFLOAT,
    api.ResourceType.IN_SPEC_SH1: api.Format.RGBA16_SFLOAT,
    api.ResourceType.OUT_DIFF_SH0: api.Format.RGBA16_SFLOAT,
    api.ResourceType.OUT_DIFF_SH1: api.Format.RGBA16_SFLOAT,
    api.ResourceType.OUT_SPEC_SH0: api.Format.RGBA16_SFLOAT,
    api.ResourceType.OUT_SPEC_SH1: api.Format.RGBA16_SFLOAT,
    api.ResourceType.IN_PENUMBRA: api.Format.R16_SFLOAT,
    api.ResourceType.OUT_SHADOW_TRANSLUCENCY: api.Format.R8_UNORM,   # RGBA8_UNORM for SIGMA_SHADOW_TRANSLUCENCY ( pass fmt= explicitly )
    api.ResourceType.IN_TRANSLUCENCY: api.Format.RGBA8_UNORM,
    api.ResourceType.OUT_VALIDATION: api.Format.RGBA8_UNORM,   # "RGBA8+", .w = transparency ( NRDDescs.h:142-144 )
    api.ResourceType.IN_SIGNAL: api.Format.RGBA16_SFLOAT,    # NRDSample's "Composed" ( Source/NRDSample.cpp:484-485, 2944 )
    api.ResourceType.OUT_SIGNAL: api.Format.RGBA16_SFLOAT,
    # the application's choice ( Texture2D<float> in the shaders ); these are what synth.reblur_frame( guides=True ) makes
    api.ResourceType.IN_DIFF_CONFIDENCE: api.Format.RGBA16_SFLOAT,
    api.ResourceType.IN_SPEC_CONFIDENCE: api.Format.R8_UNORM,
    api.ResourceType.IN_DISOCCLUSION_THRESHOLD_MIX: api.Format.R16_SFLOAT,
}


def alloc_texture(fmt: int, width: int, height: int, device="cpu") -> torch.Tensor:
    dtype, ch = FORMAT_STORAGE[api.Format(fmt)]
    shape = (height, width) if ch == 1 else (height, width, ch)
    return torch.zeros(shape, dtype=dtype, device=device)


def tex_desc(t: torch.Tensor, fmt: int) -> OracleTexture:
    assert t.is_contiguous() and t.device.type == "cpu"
    h, w = t.shape[0], t.shape[1]
    return OracleTexture(t.data_ptr(), w, h, w * api.FORMAT_BYTES[api.Format(fmt)], int(fmt))


class OracleDenoiser:
    """nrd::Instance + pools + oracle replay for ONE denoiser on host memory."""

    def __init__(self, host_lib: api.NrdLibrary, denoiser: int, width: int, height: int, identifier: int = 0, quads: bool = True, robust_mirror_test: bool = False,
                 engine: str = "oracle"):
        """engine: "oracle" = the hand-written restatement (liboracle.so); "reference" = the reference's shaders compiled as C++."""
        self.engine = engine
        self.width, self.height, self.identifier, self.denoiser = width, height, identifier, denoiser
        self.instance = api.NrdInstance(host_lib, [(identifier, denoiser)])
        assert self.instance.result == api.Result.SUCCESS, self.instance.result
        perm, tran = self.instance.pools()
        self.formats: Dict[tuple, int] = {}
        self.textures: Dict[tuple, torch.Tensor] = {}
        for kind, pool in ((api.ResourceType.PERMANENT_POOL, perm), (api.ResourceType.TRANSIENT_POOL, tran)):
            for i, (fmt, ds) in enumerate(pool):
                w, h = (width + ds - 1) // ds, (height + ds - 1) // ds
                self.textures[(int(kind), i)] = alloc_texture(fmt, w, h)
                self.formats[(int(kind), i)] = fmt
        self.flags = (1 if quads else 0) | (2 if robust_mirror_test else 0)
        self.last_dispatches: List[api.Dispatch] = []

    def set_user_texture(self, rtype: int, tensor: torch.Tensor, fmt: Optional[int] = None):
        self.textures[(int(rtype), 0)] = tensor
        self.formats[(int(rtype), 0)] = int(fmt if fmt is not None else USER_FORMATS[api.ResourceType(rtype)])

    def _resolve(self, b: api.Binding):
        key = (b.type, b.index) if b.type in (int(api.ResourceType.PERMANENT_POOL), int(api.ResourceType.TRANSIENT_POOL)) else (b.type, 0)
        return key

    def denoise(self, common: api.CommonSettings, settings=None,
                on_dispatch: Optional[Callable[[int, api.Dispatch, List[tuple], "OracleDenoiser"], None]] = None,
                before_dispatch: Optional[Callable[[int, api.Dispatch, List[tuple], "OracleDenoiser"], None]] = None):
        r = self.instance.set_common_settings(common)
        assert r == api.Result.SUCCESS, r
        if settings is not None:
            assert self.instance.set_denoiser_settings(self.identifier, settings) == api.Result.SUCCESS
        r, dispatches = self.instance.get_compute_dispatches([self.identifier])
        assert r == api.Result.SUCCESS, r
        self.last_dispatches = dispatches
        if self.engine == "reference":
            fn = ref_shaders().nrd_refshader_dispatch
        else:
            fn = lib().nrd_oracle_dispatch
        for i, d in enumerate(dispatches):
            keys = [self._resolve(b) for b in d.bindings]
            if before_dispatch:
                before_dispatch(i, d, keys, self)
            arr = (OracleTexture * len(keys))(*[tex_desc(self.textures[k], self.formats[k]) for k in keys])
            cb = C.create_string_buffer(d.constants, len(d.constants)) if d.constants else None
            rc = fn(d.shader.encode(), cb, len(d.constants), arr, len(keys), d.grid[0], d.grid[1], self.flags)
            assert rc == 0, f"{self.engine} dispatch failed rc={rc} for {d.shader}"
            if on_dispatch:
                on_dispatch(i, d, keys, self)
        return dispatches


def default_host_library() -> api.NrdLibrary:
    """The host library under test drives the oracle by default; tests that want the reference's own dispatch
    stream pass api.NrdLibrary(REF_LIB_PATH) instead."""
    from nrd_sample_b200 import build as b
    return api.NrdLibrary(b.host_library_path())
