// REFERENCE denoiser on sm_100a: plain temporal accumulation of IN_SIGNAL into an RGBA32F history and the copy to OUT_SIGNAL.
// What it replaces: External/NRD/Shaders/REFERENCE_TemporalAccumulation.cs.hlsl:18-28 and REFERENCE_Copy.cs.hlsl:18-27.
// Both are one-texel-in / one-texel-out streaming passes: a warp owns 32 consecutive texels of a row (512-byte history lines).
// The signal textures belong to the application (Texture2D<float4> in the shaders): RGBA16F (NRDSample's "Composed"), RGBA32F and
// RGBA8 are accepted.
#include <string>

#include "../../../include/nrd_b200.h"
#include "../../../include/nrdcu.h"
#include "../pipeline_key.h"
#include "common.cuh"

namespace nrdk {

namespace {

using nrdb::ReferenceAccumulateConstants;
using nrdb::ReferenceCopyConstants;

struct TexAny4 : TexView {
    enum Kind : uint32_t { RGBA8 = 0, RGBA16F = 1, RGBA32F = 2 };
    uint32_t kind;
    NRD_DEV float4 load(int x, int y) const {
        if (!inside(x, y)) return f4(0.0f);
        switch (kind) {
            case RGBA8: {
                uchar4 v = __ldg(ptr<uchar4>(x, y));
                return make_float4((float)v.x / 255.0f, (float)v.y / 255.0f, (float)v.z / 255.0f, (float)v.w / 255.0f);
            }
            case RGBA16F: return TexRGBA16F::decode(__ldg(ptr<uint2>(x, y)));
            default: return __ldg(ptr<float4>(x, y));
        }
    }
    NRD_DEV void store(int x, int y, float4 v) const {
        if (!inside(x, y)) return;
        switch (kind) {
            case RGBA8:
                *ptrw<uchar4>(x, y) = make_uchar4((unsigned char)unormQ(v.x, 255.0f), (unsigned char)unormQ(v.y, 255.0f), (unsigned char)unormQ(v.z, 255.0f), (unsigned char)unormQ(v.w, 255.0f));
                break;
            case RGBA16F: {
                __half2 lo = __floats2half2_rn(v.x, v.y), hi = __floats2half2_rn(v.z, v.w);
                uint2 raw;
                raw.x = *reinterpret_cast<uint32_t*>(&lo);
                raw.y = *reinterpret_cast<uint32_t*>(&hi);
                *ptrw<uint2>(x, y) = raw;
                break;
            }
            default: *ptrw<float4>(x, y) = v;
        }
    }
};

__global__ void __launch_bounds__(256) referenceTemporalAccumulationKernel(const __grid_constant__ ReferenceAccumulateConstants cb, const __grid_constant__ TexAny4 input,
                                                                          const __grid_constant__ TexAny4 history) {
    pdlEntry();
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    const float4 in = input.load(x, y), h = history.load(x, y);
    history.store(x, y, lerp(h, in, cb.accumSpeed));
}

__global__ void __launch_bounds__(256) referenceCopyKernel(const __grid_constant__ ReferenceCopyConstants cb, const __grid_constant__ TexAny4 input, const __grid_constant__ TexAny4 output,
                                                          int w, int h) {
    pdlEntry();
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= w || y >= h) return;
    const float u = ((float)x + 0.5f) * cb.rectSizeInv[0];
    if (u > cb.splitScreen) output.store(x, y, input.load(x, y));
}

bool bindAny4(const nrdcuTexture& t, TexAny4& v) {
    uint32_t bpp;
    switch ((nrd::Format)t.format) {
        case nrd::Format::RGBA8_UNORM: v.kind = TexAny4::RGBA8; bpp = 4; break;
        case nrd::Format::RGBA16_SFLOAT: v.kind = TexAny4::RGBA16F; bpp = 8; break;
        case nrd::Format::RGBA32_SFLOAT: v.kind = TexAny4::RGBA32F; bpp = 16; break;
        default: return false;
    }
    if (!t.data || (t.pitchBytes % bpp) != 0 || t.pitchBytes < t.width * bpp || ((uintptr_t)t.data % bpp) != 0) return false;
    v.data = (uint8_t*)t.data;
    v.w = (int)t.width;
    v.h = (int)t.height;
    v.pitch = (int)(t.pitchBytes / bpp);
    v.fmt = t.format;
    return true;
}

}  // namespace

// Dispatch by shader identifier (called by the executor). `gridW` x `gridH`: the 16x16 groups of the DispatchDesc (the constants of
// the accumulation pass carry no rect size); 0 = cover the whole history texture.
uint32_t dispatchReference(const PipelineKey& key, const void* constants, uint32_t constantsSize, const nrdcuTexture* tex, uint32_t n, uint32_t gridW, uint32_t gridH,
                           cudaStream_t stream, std::string& err) {
    using nrd::Result;
    const std::string id = key.id;   // ( short: fits the small-string buffer; used by the messages only )
    TexAny4 a, b;
    if (n != 2 || !bindAny4(tex[0], a) || !bindAny4(tex[1], b)) {
        err = id + ": expects an input and an output in RGBA8 / RGBA16F / RGBA32F";
        return (uint32_t)Result::INVALID_ARGUMENT;
    }
    if (key.pass == REFERENCE_TEMPORAL_ACCUMULATION) {
        if (constantsSize != sizeof(ReferenceAccumulateConstants) || !constants) {
            err = id + ": expected 16 constant bytes";
            return (uint32_t)Result::INVALID_ARGUMENT;
        }
        ReferenceAccumulateConstants cb;
        memcpy(&cb, constants, sizeof(cb));
        const int w = gridW ? min((int)gridW * 16, b.w) : b.w, h = gridH ? min((int)gridH * 16, b.h) : b.h;
        launchK(referenceTemporalAccumulationKernel, dim3((w + 31) / 32, (h + 7) / 8), dim3(32, 8), 0, stream, cb, a, b);
    } else if (key.pass == REFERENCE_COPY) {
        if (constantsSize != sizeof(ReferenceCopyConstants) || !constants) {
            err = id + ": expected 24 constant bytes";
            return (uint32_t)Result::INVALID_ARGUMENT;
        }
        ReferenceCopyConstants cb;
        memcpy(&cb, constants, sizeof(cb));
        // the rect: 1 / gRectSizeInv (the reference dispatches ceil( rect / 16 ) groups; texels beyond the rect are never written)
        int w = min(b.w, (int)(1.0f / cb.rectSizeInv[0] + 0.5f)), h = min(b.h, (int)(1.0f / cb.rectSizeInv[1] + 0.5f));
        if (gridW) w = min(w, (int)gridW * 16);
        if (gridH) h = min(h, (int)gridH * 16);
        launchK(referenceCopyKernel, dim3((w + 31) / 32, (h + 7) / 8), dim3(32, 8), 0, stream, cb, a, b, w, h);
    } else {
        err = "no CUDA kernel for shader '" + id + "'";
        return (uint32_t)Result::UNSUPPORTED;
    }
    return (uint32_t)Result::SUCCESS;
}

}  // namespace nrdk
