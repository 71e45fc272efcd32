"""ctypes binding of the CUDA executor C ABI (include/nrdcu.h) + a thin per-denoiser wrapper.

torch is used for device memory and streams only; every pixel is produced by the hand-written kernels in
csrc/kernels/*.cu behind `nrdcuDispatch` / `nrdcuDenoise`. If the library or a CUDA device is missing the calls
raise — there is no fallback path.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Optional, Sequence

import torch

from . import build as _build
from . import nrd_api as api


class CuTexture(C.Structure):
    _fields_ = [("data", C.c_void_p), ("width", C.c_uint32), ("height", C.c_uint32), ("pitchBytes", C.c_uint32), ("format", C.c_uint32)]


FLAG_QUAD_INTRINSICS = 1
FLAG_CUDA_GRAPH = 2
FLAG_ROBUST_MIRROR_TEST = 4
FLAG_PROBE_MIRROR = 8
FLAG_NO_SEAM_OVERLAP = 16

NRDCU_SYMBOLS = ("nrdcuDispatch", "nrdcuDispatchRows", "nrdcuDenoiseRows", "nrdcuAllocSharedTexture", "nrdcuTileExportSize", "nrdcuTileExport", "nrdcuTileAttach",
                 "nrdcuTileSetHalo", "nrdcuTileGetStatus", "nrdcuTileGetMotionBound", "nrdcuCreate", "nrdcuDestroy", "nrdcuSetCommonSettings", "nrdcuSetDenoiserSettings", "nrdcuSetResource", "nrdcuDenoise",
                 "nrdcuGetPoolTexture", "nrdcuGetInstance", "nrdcuSetHostResource", "nrdcuDenoiseHost", "nrdcuDenoiseHostPipelined", "nrdcuHostFlush", "nrdcuGetLastError", "nrdcuGetLaunchCount",
                 "nrdcuHostFrameCreate", "nrdcuHostFrameGetTexture", "nrdcuHostFrameGetInfo", "nrdcuHostFrameDestroy", "nrdcuDenoiseHostFrames",
                 "nrdcuGetPoolBytes", "nrdcuGetMemoryUsage", "nrdcuGetGraphStats", "nrdcuGetMirrorProbe", "nrdcuSetProfiling", "nrdcuResolveProfile", "nrdcuGetProfileEntry", "nrdcuResetProfile",
                 "nrdcuFrontEndPackNormalRoughness", "nrdcuFrontEndPackRadianceHitDist", "nrdcuBackEndUnpackRadiance", "nrdcuFrontEndProbe", "nrdcuFrontEndGetLastError")

# nrd::Format -> (torch dtype, channels) for tensors handed to / returned by the executor
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

# void (*nrdcuDispatchCallback)(void* userArg, uint32_t dispatchIndex, const char* passName, const nrdcuTexture*, const uint8_t* isStorage, uint32_t n)
DISPATCH_CALLBACK = C.CFUNCTYPE(None, C.c_void_p, C.c_uint32, C.c_char_p, C.POINTER(CuTexture), C.POINTER(C.c_uint8), C.c_uint32)

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        path = os.environ.get("NRD_B200_LIB") or _build.build()   # NRD_B200_LIB: a prebuilt variant (tools/build_variant.py), experiments only
        L = C.CDLL(path)
        L.nrdcuDispatch.argtypes = [C.c_char_p, C.c_void_p, C.c_uint32, C.POINTER(CuTexture), C.c_uint32, C.c_uint32, C.c_void_p]
        L.nrdcuDispatch.restype = C.c_uint32
        L.nrdcuDispatchRows.argtypes = [C.c_char_p, C.c_void_p, C.c_uint32, C.POINTER(CuTexture), C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32]
        L.nrdcuDispatchRows.restype = C.c_uint32
        L.nrdcuDenoiseRows.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32, DISPATCH_CALLBACK, C.c_void_p]
        L.nrdcuDenoiseRows.restype = C.c_uint32
        L.nrdcuAllocSharedTexture.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(CuTexture)]
        L.nrdcuAllocSharedTexture.restype = C.c_uint32
        L.nrdcuTileExportSize.argtypes = [C.c_void_p]
        L.nrdcuTileExportSize.restype = C.c_uint32
        L.nrdcuTileExport.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        L.nrdcuTileExport.restype = C.c_uint32
        L.nrdcuTileAttach.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
        L.nrdcuTileAttach.restype = C.c_uint32
        L.nrdcuTileSetHalo.argtypes = [C.c_void_p, C.c_char_p, C.c_uint32, C.c_uint32]
        L.nrdcuTileSetHalo.restype = C.c_uint32
        L.nrdcuTileGetStatus.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]
        L.nrdcuTileGetStatus.restype = C.c_uint32
        L.nrdcuTileGetMotionBound.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.nrdcuTileGetMotionBound.restype = C.c_uint32
        L.nrdcuCreate.argtypes = [C.c_void_p, C.c_uint16, C.c_uint16, C.c_int, C.c_uint32, C.POINTER(C.c_void_p)]
        L.nrdcuCreate.restype = C.c_uint32
        L.nrdcuDestroy.argtypes = [C.c_void_p]
        L.nrdcuDestroy.restype = None
        L.nrdcuSetCommonSettings.argtypes = [C.c_void_p, C.c_void_p]
        L.nrdcuSetCommonSettings.restype = C.c_uint32
        L.nrdcuSetDenoiserSettings.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        L.nrdcuSetDenoiserSettings.restype = C.c_uint32
        L.nrdcuSetResource.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(CuTexture)]
        L.nrdcuSetResource.restype = C.c_uint32
        L.nrdcuDenoise.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.c_uint32, C.c_void_p]
        L.nrdcuDenoise.restype = C.c_uint32
        L.nrdcuGetPoolTexture.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.POINTER(CuTexture)]
        L.nrdcuGetPoolTexture.restype = C.c_uint32
        L.nrdcuGetInstance.argtypes = [C.c_void_p]
        L.nrdcuGetInstance.restype = C.c_void_p
        L.nrdcuSetHostResource.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int]
        L.nrdcuSetHostResource.restype = C.c_uint32
        L.nrdcuDenoiseHost.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.c_uint32, C.c_void_p]
        L.nrdcuDenoiseHost.restype = C.c_uint32
        L.nrdcuDenoiseHostPipelined.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.c_uint32, C.c_void_p]
        L.nrdcuDenoiseHostPipelined.restype = C.c_uint32
        L.nrdcuHostFlush.argtypes = [C.c_void_p, C.c_void_p]
        L.nrdcuHostFlush.restype = C.c_uint32
        L.nrdcuGetLastError.restype = C.c_char_p
        L.nrdcuGetLaunchCount.restype = C.c_uint64
        L.nrdcuGetPoolBytes.argtypes = [C.c_void_p]
        L.nrdcuGetPoolBytes.restype = C.c_uint64
        L.nrdcuGetGraphStats.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]
        L.nrdcuGetGraphStats.restype = C.c_uint32
        L.nrdcuSetProfiling.argtypes = [C.c_void_p, C.c_int]
        L.nrdcuSetProfiling.restype = C.c_uint32
        L.nrdcuResolveProfile.argtypes = [C.c_void_p]
        L.nrdcuResolveProfile.restype = C.c_uint32
        L.nrdcuGetProfileEntry.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_char_p), C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
        L.nrdcuGetProfileEntry.restype = C.c_uint32
        L.nrdcuResetProfile.argtypes = [C.c_void_p]
        L.nrdcuResetProfile.restype = None
        L.nrdcuHostFrameCreate.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
        L.nrdcuHostFrameCreate.restype = C.c_uint32
        L.nrdcuHostFrameGetTexture.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(CuTexture)]
        L.nrdcuHostFrameGetTexture.restype = C.c_uint32
        L.nrdcuHostFrameGetInfo.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_int)]
        L.nrdcuHostFrameGetInfo.restype = C.c_uint32
        L.nrdcuHostFrameDestroy.argtypes = [C.c_void_p]
        L.nrdcuHostFrameDestroy.restype = None
        L.nrdcuDenoiseHostFrames.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.nrdcuDenoiseHostFrames.restype = C.c_uint32
        L.nrdcuGetMirrorProbe.argtypes = [C.POINTER(C.c_uint64), C.c_int]
        L.nrdcuGetMirrorProbe.restype = C.c_uint32
        _lib = L
    return _lib


class NrdcuError(RuntimeError):
    pass


def _check(rc: int, what: str):
    if rc != 0:
        raise NrdcuError(f"{what}: {api.Result(rc).name}: {load().nrdcuGetLastError().decode()}")


def launch_count() -> int:
    return int(load().nrdcuGetLaunchCount())


def mirror_probe(reset: bool = False):
    """(taps, taps that took the "mirrored" weight branch) counted by the spatial passes of contexts created with FLAG_PROBE_MIRROR."""
    out = (C.c_uint64 * 14)()
    _check(load().nrdcuGetMirrorProbe(out, 1 if reset else 0), "nrdcuGetMirrorProbe")
    return int(out[0]), int(out[1])


def mirror_probe_detail():
    """{(pass, lobe): (taps, mirrored)} of the same counters; pass in ("Pre-pass", "Blur", "Post-blur"), lobe in ("diff", "spec")."""
    out = (C.c_uint64 * 14)()
    _check(load().nrdcuGetMirrorProbe(out, 0), "nrdcuGetMirrorProbe")
    return {(p, l): (int(out[2 + 2 * (i * 2 + j)]), int(out[3 + 2 * (i * 2 + j)])) for i, p in enumerate(("Pre-pass", "Blur", "Post-blur")) for j, l in enumerate(("diff", "spec"))}


def texture_of(t: torch.Tensor, fmt: int) -> CuTexture:
    """View a contiguous tensor (H, W[, C]) as an nrdcuTexture. The tensor must stay alive while the texture is in use."""
    assert t.is_contiguous(), "textures are pitch-linear: tensor must be contiguous"
    h, w = t.shape[0], t.shape[1]
    return CuTexture(t.data_ptr(), w, h, w * api.FORMAT_BYTES[api.Format(fmt)], int(fmt))


def alloc_texture(fmt: int, width: int, height: int, device) -> torch.Tensor:
    dtype, ch = FORMAT_STORAGE[api.Format(fmt)]
    # keep rows 16-byte aligned: width * bytes is a multiple of 16 for every width that is a multiple of 16 / bpp
    shape = (height, width) if ch == 1 else (height, width, ch)
    return torch.zeros(shape, dtype=dtype, device=device)


def dispatch(shader: str, constants: bytes, textures: Sequence[CuTexture], flags: int = FLAG_QUAD_INTRINSICS, stream: Optional[torch.cuda.Stream] = None,
             rows: Optional[Sequence[int]] = None):
    """One pass (== one nrd::DispatchDesc) on caller-provided device textures; `rows` = (begin, end) restricts it to a strip."""
    arr = (CuTexture * len(textures))(*textures)
    cb = C.create_string_buffer(constants, len(constants)) if constants else None
    s = stream.cuda_stream if stream is not None else torch.cuda.current_stream().cuda_stream
    if rows is None:
        _check(load().nrdcuDispatch(shader.encode(), cb, len(constants), arr, len(textures), flags, C.c_void_p(s)), f"nrdcuDispatch({shader})")
    else:
        _check(load().nrdcuDispatchRows(shader.encode(), cb, len(constants), arr, len(textures), flags, C.c_void_p(s), rows[0], rows[1]), f"nrdcuDispatchRows({shader})")


class CudaDenoiser:
    """nrdcuContext wrapper for ONE denoiser: pools live in the C++ context, user textures are torch CUDA tensors."""

    def __init__(self, denoiser: int, width: int, height: int, identifier: int = 0, device: int = 0, flags: int = FLAG_QUAD_INTRINSICS):
        self.width, self.height, self.identifier, self.denoiser, self.device = width, height, identifier, denoiser, device
        L = load()
        self._denoisers = (api.DenoiserDesc * 1)(api.DenoiserDesc(identifier, int(denoiser)))
        desc = api.InstanceCreationDesc()
        desc.denoisers = self._denoisers
        desc.denoisersNum = 1
        ctx = C.c_void_p()
        _check(L.nrdcuCreate(C.byref(desc), width, height, device, flags, C.byref(ctx)), "nrdcuCreate")
        self.ctx = ctx
        self._keep: Dict[int, torch.Tensor] = {}
        self._ids = (C.c_uint32 * 1)(identifier)

    def close(self):
        if getattr(self, "ctx", None):
            load().nrdcuDestroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_user_texture(self, rtype: int, tensor: torch.Tensor, fmt: int):
        assert tensor.is_cuda
        tex = texture_of(tensor, fmt)
        _check(load().nrdcuSetResource(self.ctx, int(rtype), C.byref(tex)), "nrdcuSetResource")
        self._keep[int(rtype)] = tensor

    def set_host_texture(self, rtype: int, tensor: torch.Tensor, fmt: int, is_output: bool):
        """Plugin-style path: host buffer in, staging + copies inside nrdcuDenoiseHost."""
        assert not tensor.is_cuda and tensor.is_contiguous()
        h, w = tensor.shape[0], tensor.shape[1]
        _check(load().nrdcuSetHostResource(self.ctx, int(rtype), C.c_void_p(tensor.data_ptr()), w, h, w * api.FORMAT_BYTES[api.Format(fmt)], int(fmt), 1 if is_output else 0),
               "nrdcuSetHostResource")
        self._keep[1000 + int(rtype)] = tensor

    def set_common_settings(self, cs: api.CommonSettings):
        _check(load().nrdcuSetCommonSettings(self.ctx, C.byref(cs)), "nrdcuSetCommonSettings")

    def set_denoiser_settings(self, settings):
        _check(load().nrdcuSetDenoiserSettings(self.ctx, self.identifier, C.byref(settings)), "nrdcuSetDenoiserSettings")

    def denoise(self, stream: Optional[torch.cuda.Stream] = None):
        s = stream.cuda_stream if stream is not None else torch.cuda.current_stream().cuda_stream
        _check(load().nrdcuDenoise(self.ctx, self._ids, 1, C.c_void_p(s)), "nrdcuDenoise")

    def denoise_host(self, stream: Optional[torch.cuda.Stream] = None):
        s = stream.cuda_stream if stream is not None else torch.cuda.current_stream().cuda_stream
        _check(load().nrdcuDenoiseHost(self.ctx, self._ids, 1, C.c_void_p(s)), "nrdcuDenoiseHost")

    def denoise_host_pipelined(self, stream: Optional[torch.cuda.Stream] = None):
        """Host buffers in / out with the PCIe copies of neighbouring frames overlapping the kernels; call host_flush() before reading outputs."""
        s = stream.cuda_stream if stream is not None else torch.cuda.current_stream().cuda_stream
        _check(load().nrdcuDenoiseHostPipelined(self.ctx, self._ids, 1, C.c_void_p(s)), "nrdcuDenoiseHostPipelined")

    def host_flush(self, stream: Optional[torch.cuda.Stream] = None):
        s = stream.cuda_stream if stream is not None else torch.cuda.current_stream().cuda_stream
        _check(load().nrdcuHostFlush(self.ctx, C.c_void_p(s)), "nrdcuHostFlush")

    # ---- host frames: one pinned, NUMA-local host block per direction, one copy up and one down per call ------------------------------------
    def declare_host_texture(self, rtype: int, fmt: int, is_output: bool, width: Optional[int] = None, height: Optional[int] = None):
        """Make `rtype` part of the host frames of its direction ( call for every host resource before the first host_frame() )."""
        _check(load().nrdcuSetHostResource(self.ctx, int(rtype), None, width or self.width, height or self.height, 0, int(fmt), 1 if is_output else 0), "nrdcuSetHostResource")

    def host_frame(self, is_output: bool) -> "HostFrame":
        return HostFrame(self, is_output)

    def denoise_host_frames(self, inputs: "HostFrame", outputs: "HostFrame", stream: Optional[torch.cuda.Stream] = None):
        """Upload `inputs` ( one copy ), run the chain, download into `outputs` ( one copy ), pipelined across calls; host_flush() before reading."""
        s = stream.cuda_stream if stream is not None else torch.cuda.current_stream().cuda_stream
        _check(load().nrdcuDenoiseHostFrames(self.ctx, self._ids, 1, inputs.handle, outputs.handle, C.c_void_p(s)), "nrdcuDenoiseHostFrames")

    def pool_texture(self, permanent: bool, index: int) -> torch.Tensor:
        """Copy of a pool texture (debug / parity tap)."""
        tex = CuTexture()
        _check(load().nrdcuGetPoolTexture(self.ctx, 1 if permanent else 0, index, C.byref(tex)), "nrdcuGetPoolTexture")
        dtype, ch = FORMAT_STORAGE[api.Format(tex.format)]
        out = alloc_texture(tex.format, tex.width, tex.height, f"cuda:{self.device}")
        row = tex.width * api.FORMAT_BYTES[api.Format(tex.format)]
        # cudaMemcpy2D through torch: wrap the pitched allocation as a byte tensor view
        src = _as_byte_tensor(tex.data, tex.pitchBytes * tex.height, self.device).view(tex.height, tex.pitchBytes)[:, :row]
        out.view(torch.uint8).view(tex.height, row).copy_(src)
        return out

    def set_profiling(self, enabled: bool):
        _check(load().nrdcuSetProfiling(self.ctx, 1 if enabled else 0), "nrdcuSetProfiling")

    def reset_profile(self):
        load().nrdcuResetProfile(self.ctx)

    def profile(self) -> Dict[str, tuple]:
        """{pass name: (total ms, launches)} accumulated since the last reset (synchronises the device)."""
        n = load().nrdcuResolveProfile(self.ctx)
        out = {}
        for i in range(n):
            name, ms, cnt = C.c_char_p(), C.c_double(), C.c_uint64()
            _check(load().nrdcuGetProfileEntry(self.ctx, i, C.byref(name), C.byref(ms), C.byref(cnt)), "nrdcuGetProfileEntry")
            out[name.value.decode()] = (ms.value, cnt.value)
        return out

    def pool_bytes(self) -> int:
        return int(load().nrdcuGetPoolBytes(self.ctx))

    def graph_stats(self) -> dict:
        """FLAG_CUDA_GRAPH bookkeeping: frames captured into a new graph, frames replayed from a cached one, graphs cached."""
        cap, rep, cached = C.c_uint64(0), C.c_uint64(0), C.c_uint32(0)
        _check(load().nrdcuGetGraphStats(self.ctx, C.byref(cap), C.byref(rep), C.byref(cached)), "nrdcuGetGraphStats")
        return {"captures": cap.value, "replays": rep.value, "cached": cached.value}


class HostFrame:
    """nrdcuHostFrame: the renderer-facing staging block. `tensor( rtype )` is a CPU tensor VIEW ( H, W[, C] ) of the resource inside the pinned
    block ( rows are 256-byte aligned: the view may be strided )."""

    def __init__(self, den: CudaDenoiser, is_output: bool):
        self.den, self.is_output = den, is_output
        h = C.c_void_p()
        _check(load().nrdcuHostFrameCreate(den.ctx, 1 if is_output else 0, C.byref(h)), "nrdcuHostFrameCreate")
        self.handle = h
        b, n = C.c_uint64(), C.c_int()
        _check(load().nrdcuHostFrameGetInfo(h, C.byref(b), C.byref(n)), "nrdcuHostFrameGetInfo")
        self.bytes, self.numa_node = int(b.value), int(n.value)

    def tensor(self, rtype: int) -> torch.Tensor:
        tex = CuTexture()
        _check(load().nrdcuHostFrameGetTexture(self.handle, int(rtype), C.byref(tex)), "nrdcuHostFrameGetTexture")
        dtype, ch = FORMAT_STORAGE[api.Format(tex.format)]
        raw = torch.frombuffer((C.c_uint8 * (tex.pitchBytes * tex.height)).from_address(tex.data), dtype=torch.uint8).view(dtype)
        pitch = tex.pitchBytes // raw.element_size()
        return raw.as_strided((tex.height, tex.width, ch), (pitch, ch, 1)) if ch > 1 else raw.as_strided((tex.height, tex.width), (pitch, 1))

    def close(self):
        if self.handle:
            load().nrdcuHostFrameDestroy(self.handle)
            self.handle = None


def _as_byte_tensor(ptr: int, nbytes: int, device: int) -> torch.Tensor:
    """Zero-copy uint8 view of raw device memory (via __cuda_array_interface__)."""

    class _Raw:
        pass

    r = _Raw()
    r.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 3}
    return torch.as_tensor(r, device=f"cuda:{device}")


# ------------------------------------------------------------------------------------------------
# Front end / back end on the device (include/nrd_frontend.cuh, csrc/kernels/frontend.cu)
# ------------------------------------------------------------------------------------------------
def _frontend_check(rc: int, what: str):
    if rc != 0:
        raise NrdcuError(f"{what}: {api.Result(rc).name}: {load().nrdcuFrontEndGetLastError().decode()}")


def frontend_probe(inputs: Sequence[torch.Tensor], device="cuda:0"):
    """Runs csrc/frontend_probe.inl on the device: `inputs` = 6 (n, 4) fp32 tensors, returns 21 (n, 4) fp32 device tensors."""
    L = load()
    L.nrdcuFrontEndProbe.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_uint32, C.c_void_p]
    L.nrdcuFrontEndGetLastError.restype = C.c_char_p
    n = inputs[0].shape[0]
    dev_in = [t.to(device=device, dtype=torch.float32).contiguous() for t in inputs]
    outs = [torch.zeros(n, 4, dtype=torch.float32, device=device) for _ in range(21)]
    torch.cuda.synchronize()
    rc = L.nrdcuFrontEndProbe((C.c_void_p * 6)(*[t.data_ptr() for t in dev_in]), (C.c_void_p * 21)(*[t.data_ptr() for t in outs]), n, None)
    _frontend_check(rc, "nrdcuFrontEndProbe")
    return outs


def frontend_pack_normal_roughness(normal_roughness: torch.Tensor, material_id: Optional[torch.Tensor] = None) -> torch.Tensor:
    """(H, W, 4) fp32 { N, linear roughness } (+ (H, W) fp32 material IDs 0..3) -> IN_NORMAL_ROUGHNESS (H, W) int32 texels."""
    L = load()
    L.nrdcuFrontEndPackNormalRoughness.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(CuTexture), C.c_void_p]
    L.nrdcuFrontEndGetLastError.restype = C.c_char_p
    h, w = normal_roughness.shape[:2]
    out = alloc_texture(api.Format.R10_G10_B10_A2_UNORM, w, h, normal_roughness.device)
    tex = texture_of(out, api.Format.R10_G10_B10_A2_UNORM)
    rc = L.nrdcuFrontEndPackNormalRoughness(normal_roughness.contiguous().data_ptr(), material_id.contiguous().data_ptr() if material_id is not None else None, C.byref(tex),
                                            C.c_void_p(torch.cuda.current_stream().cuda_stream))
    _frontend_check(rc, "nrdcuFrontEndPackNormalRoughness")
    return out


def frontend_pack_radiance_hitdist(mode: int, radiance_hitdist: torch.Tensor, viewz: Optional[torch.Tensor] = None, normal_roughness: Optional[torch.Tensor] = None,
                                   is_specular: bool = False, hit_dist_params=(3.0, 0.1, 20.0)) -> torch.Tensor:
    """(H, W, 4) fp32 { linear radiance, hit distance } -> IN_*_RADIANCE_HITDIST (H, W, 4) fp16. mode 0 = REBLUR (needs viewz, and normal_roughness for specular), 1 = RELAX."""
    L = load()
    L.nrdcuFrontEndPackRadianceHitDist.argtypes = [C.c_uint32, C.c_void_p, C.POINTER(CuTexture), C.POINTER(CuTexture), C.c_uint32, C.POINTER(C.c_float), C.POINTER(CuTexture), C.c_void_p]
    L.nrdcuFrontEndGetLastError.restype = C.c_char_p
    h, w = radiance_hitdist.shape[:2]
    out = alloc_texture(api.Format.RGBA16_SFLOAT, w, h, radiance_hitdist.device)
    tz = C.byref(texture_of(viewz, api.Format.R32_SFLOAT)) if viewz is not None else None
    tn = C.byref(texture_of(normal_roughness, api.Format.R10_G10_B10_A2_UNORM)) if normal_roughness is not None else None
    to = texture_of(out, api.Format.RGBA16_SFLOAT)
    rc = L.nrdcuFrontEndPackRadianceHitDist(mode, radiance_hitdist.contiguous().data_ptr(), tz, tn, 1 if is_specular else 0, (C.c_float * 3)(*hit_dist_params), C.byref(to),
                                            C.c_void_p(torch.cuda.current_stream().cuda_stream))
    _frontend_check(rc, "nrdcuFrontEndPackRadianceHitDist")
    return out


def backend_unpack_radiance(mode: int, texture: torch.Tensor) -> torch.Tensor:
    """OUT_*_RADIANCE_HITDIST (H, W, 4) fp16 -> (H, W, 4) fp32 { linear radiance, .w as stored }."""
    L = load()
    L.nrdcuBackEndUnpackRadiance.argtypes = [C.c_uint32, C.POINTER(CuTexture), C.c_void_p, C.c_void_p]
    L.nrdcuFrontEndGetLastError.restype = C.c_char_p
    h, w = texture.shape[:2]
    out = torch.empty(h, w, 4, dtype=torch.float32, device=texture.device)
    tex = texture_of(texture, api.Format.RGBA16_SFLOAT)
    rc = L.nrdcuBackEndUnpackRadiance(mode, C.byref(tex), out.data_ptr(), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    _frontend_check(rc, "nrdcuBackEndUnpackRadiance")
    return out


def host_library():
    """nrd_api.NrdLibrary over the product library (the nrd:: entry points live in the same .so as the executor)."""
    from . import nrd_api
    return nrd_api.NrdLibrary(_build.build() if not os.environ.get("NRD_B200_LIB") else os.environ["NRD_B200_LIB"])
