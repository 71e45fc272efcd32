// Shared device code of the denoiser kernels: pitch-linear typed texture views with the D3D-style access
// semantics the reference shaders rely on, and the math / weight helpers the three REBLUR spatial passes, temporal
// accumulation, history fix and temporal stabilisation have in common.
//
// What each helper reproduces (file:line in /root/reference):
//   MathLib        External/NRIFramework/External/MathLib/ml.hlsli  (Math 105-400, Geometry 477-672, Packing 1179-1215,
//                  Filtering 1287-1414, Sequence 1620-1713, Rng 1752-1862, ImportanceSampling 2347-2405)
//   NRD.hlsli      External/NRD/Shaders/NRD.hlsli:361-428, 559-573, 637-684
//   Common.hlsli   External/NRD/Shaders/Common.hlsli:207-218, 253-277, 307-330, 346-364, 421-600
// Texture semantics: Load out of bounds -> 0, store out of bounds dropped, samplers clamp to edge,
// bilinear weights exact fp32, UNORM stores round to nearest, FP16 stores round to nearest even.
#pragma once
#include <cstdlib>
#include "../../../include/nrd_b200.h"
#include "../host/constants.h"
#include "vecmath.cuh"

#include <tuple>
#include <type_traits>
#include <utility>

namespace nrdk {

using nrdb::Mat4;

constexpr float NRD_EPS = 1e-6f;
constexpr float NRD_INF = 1e6f;
constexpr float NRD_NORMAL_ENCODING_ERROR = 0.75f / 255.0f;
constexpr float NRD_ROUGHNESS_SENSITIVITY = 0.01f;
constexpr float NRD_MAX_PERCENT_OF_LOBE_VOLUME = 0.75f;
constexpr float ML_SMALL_EPS = 1e-15f;
constexpr float ML_EPS = 1e-6f;

// CTA order (Common.hlsli:125-136, NRD_CTA_ORDER_REVERSED / _DEFAULT): consecutive passes walk the tile grid in opposite
// directions, so a pass starts on the tiles its producer wrote last and still finds them in the 126 MB L2.
// NRD_CTA_REV_MASK bits: 0 PrePass, 1 TemporalAccumulation, 2 HistoryFix, 3 Blur, 4 PostBlur, 5 TemporalStabilization.
#ifndef NRD_CTA_REV_MASK
#define NRD_CTA_REV_MASK 0x35  // the reference's pattern: PrePass, HistoryFix, PostBlur, TemporalStabilization reversed
#endif
// `ctaY0`: first CTA row of the launch when only a strip of rows is computed (multi-GPU strips, nrdcuDenoiseRows)
template <int BIT> NRD_DEV int2 ctaTile(int ctaY0) {
    if ((NRD_CTA_REV_MASK >> BIT) & 1) return make_int2((int)(gridDim.x - 1u - blockIdx.x), ctaY0 + (int)(gridDim.y - 1u - blockIdx.y));
    return make_int2((int)blockIdx.x, ctaY0 + (int)blockIdx.y);
}

// Rows [begin, end) of the rect a launch covers; begin must be a multiple of 16 (tile- and CTA-aligned)
struct Rows {
    int begin = 0, end = 0x7FFFFFFF;
};
struct RowGrid {
    int ctaY0;
    unsigned count;
};
inline RowGrid rowGrid(Rows r, int rectHeight, int blockHeight) {
    const int b = r.begin < 0 ? 0 : r.begin, e = r.end < rectHeight ? r.end : rectHeight;
    if (e <= b) return {0, 0u};
    return {b / blockHeight, (unsigned)((e - b + blockHeight - 1) / blockHeight)};
}

// =================================================================================================================
// Kernel launches. Every kernel of the denoiser chains goes through launchK, which does one of two things:
//  - normally: kernel<<< grid, block, smem, stream >>>( args... );
//  - while the executor REPLAYS a frame it holds as a CUDA graph ( NRDCU_FLAG_CUDA_GRAPH ): the launch that would have happened becomes the new parameter
//    set of the graph's next kernel node ( constants and views change every frame, the chain of kernels does not ). A launch that does not match the
//    recorded chain — another kernel, or more launches than nodes — marks the replay as failed and the executor captures the frame afresh.
// While a frame is being captured the launched functions are also recorded, so that a later replay can tell whether it still is the same chain.
// =================================================================================================================
struct GraphReplay {
    cudaGraphExec_t exec = nullptr;
    const cudaGraphNode_t* nodes = nullptr;
    const void* const* funcs = nullptr;
    uint32_t count = 0, cursor = 0;
    bool failed = false;
};
struct GraphRecord {
    static constexpr uint32_t kMax = 64;
    const void* funcs[kMax];
    uint32_t count = 0;
    bool overflow = false;
};
extern thread_local GraphReplay* g_graphReplay;   // executor.cu
extern thread_local GraphRecord* g_graphRecord;
extern thread_local bool g_pdlAllowed;

#ifdef __CUDACC__
// First statement of every kernel launched through launchK ( tests/test_abi.py checks the sources ): let the next kernel of the stream be scheduled as this grid's
// CTAs retire ( griddepcontrol.launch_dependents ), then wait until the previous kernel has completed and its writes are visible ( griddepcontrol.wait ). Both are
// no-ops for a launch without the programmatic-serialization attribute.
#ifndef NRD_B200_PDL_DEFAULT
#define NRD_B200_PDL_DEFAULT true
#endif
__device__ __forceinline__ void pdlEntry() {
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
template <class Tuple, size_t... I> inline void launchParamPointers(Tuple& values, void** out, std::index_sequence<I...>) { ((out[I] = (void*)&std::get<I>(values)), ...); }

template <class... P, class... A> inline void launchK(void (*kernel)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, const A&... args) {
    static_assert(sizeof...(P) == sizeof...(A), "launchK: argument count does not match the kernel's parameter list");
    GraphReplay* replay = g_graphReplay;
    if (!replay) {
        if (GraphRecord* rec = g_graphRecord) {
            if (rec->count < GraphRecord::kMax) rec->funcs[rec->count++] = (const void*)kernel;
            else rec->overflow = true;
        }
        // Programmatic dependent launch: a chain kernel may be SCHEDULED while its predecessor in the stream is still draining ( every chain kernel starts with
        // pdlEntry( ), which lets its own successor in and then blocks until the predecessor has completed and flushed — so the launch latency and the ramp-down
        // of one pass overlap the ramp-up of the next, nothing else ). Not while a frame is captured into a CUDA graph: the executor walks the captured chain
        // through plain dependency edges, and a replayed graph has no launch gaps to hide. Only whole frames of a stand-alone context ( g_pdlAllowed, set by the executor ): strips of a tiled frame and nrdcuDispatch keep plain serialization. NRD_B200_PDL=0 switches it off ( A/B of the benchmarks ).
        static const bool pdl = getenv("NRD_B200_PDL") ? getenv("NRD_B200_PDL")[0] != '0' : NRD_B200_PDL_DEFAULT;
        if (pdl && g_pdlAllowed && !g_graphRecord) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = grid;
            cfg.blockDim = block;
            cfg.dynamicSmemBytes = smem;
            cfg.stream = stream;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            (void)cudaLaunchKernelEx(&cfg, kernel, static_cast<P>(args)...);
            return;
        }
        kernel<<<grid, block, smem, stream>>>(args...);
        return;
    }
    if (replay->failed) return;
    if (replay->cursor >= replay->count || replay->funcs[replay->cursor] != (const void*)kernel) {
        replay->failed = true;
        return;
    }
    std::tuple<std::remove_cv_t<std::remove_reference_t<P>>...> values(args...);   // the kernel's own parameter types, by value
    void* pointers[sizeof...(P)];
    launchParamPointers(values, pointers, std::index_sequence_for<P...>());
    cudaKernelNodeParams np = {};
    np.func = (void*)kernel;
    np.gridDim = grid;
    np.blockDim = block;
    np.sharedMemBytes = (unsigned)smem;
    np.kernelParams = pointers;
    if (cudaGraphExecKernelNodeSetParams(replay->exec, replay->nodes[replay->cursor], &np) != cudaSuccess) {
        (void)cudaGetLastError();
        replay->failed = true;
        return;
    }
    replay->cursor++;
}
#endif

// =================================================================================================================
// Texture views. One struct per storage format so every access compiles to a single fixed-width LDG/STG.
// =================================================================================================================
struct TexView {
    uint8_t* data;
    int w, h, pitch;  // pitch in TEXELS (the executor divides the byte pitch by the texel size): address = base + (y * pitch + x) * sizeof(T)
    uint32_t fmt;     // nrd::Format of the bound texture ( sits in the struct's tail padding ): only the format-polymorphic accessors of the occlusion modes read it
    NRD_DEV bool inside(int x, int y) const { return (unsigned)x < (unsigned)w && (unsigned)y < (unsigned)h; }
    // one 32-bit IMAD for the texel index + one IMAD.WIDE for the address (textures are < 2^31 texels)
    template <class T> NRD_DEV const T* ptr(int x, int y) const { return reinterpret_cast<const T*>(data) + (y * pitch + x); }
    template <class T> NRD_DEV T* ptrw(int x, int y) const { return reinterpret_cast<T*>(data) + (y * pitch + x); }
    NRD_DEV int cx(int x) const { return clampi(x, 0, w - 1); }
    NRD_DEV int cy(int y) const { return clampi(y, 0, h - 1); }
};

NRD_DEV uint32_t unormQ(float v, float maxv) { return (uint32_t)(saturate(v) * maxv + 0.5f); }

struct TexR32F : TexView {
    NRD_DEV float fetch(int x, int y) const { return __ldg(ptr<float>(x, y)); }
    NRD_DEV float load(int x, int y) const { return inside(x, y) ? fetch(x, y) : 0.0f; }
    NRD_DEV float fetchClamped(int x, int y) const { return fetch(cx(x), cy(y)); }
    NRD_DEV void store(int x, int y, float v) const { if (inside(x, y)) *ptrw<float>(x, y) = v; }
    NRD_DEV float sampleNearest(float2 uv) const { return fetchClamped((int)floorf(uv.x * (float)w), (int)floorf(uv.y * (float)h)); }
    NRD_DEV float sampleLinear(float2 uv) const {
        float tx = uv.x * (float)w - 0.5f, ty = uv.y * (float)h - 0.5f;
        float fx = floorf(tx), fy = floorf(ty);
        float wx = tx - fx, wy = ty - fy;
        int x0 = (int)fx, y0 = (int)fy;
        float a = fetchClamped(x0, y0), b = fetchClamped(x0 + 1, y0), c = fetchClamped(x0, y0 + 1), d = fetchClamped(x0 + 1, y0 + 1);
        return lerp(lerp(a, b, wx), lerp(c, d, wx), wy);
    }
};

// Sky-tile classification ( *_ClassifyTiles.cs.hlsl ): is every pixel of the 16x16 tile ( tx, ty ) outside the denoising range? One WARP per tile: each lane takes 8
// consecutive pixels of a row as two 128-bit loads ( rows 16-byte aligned and the tile inside the texture; pixel by pixel with out-of-bounds = 0 otherwise ), the
// verdict is one warp vote — no shared memory, no barrier, 8 tiles per 256-thread CTA. `isSky( rawViewZ )` is the denoiser's own range test.
template <class IsSky> NRD_DEV bool tileIsSkyWarp(const TexR32F& viewZ, int tx, int ty, IsSky isSky) {
    const int lane = threadIdx.x & 31;
    const int px0 = tx * 16 + (lane & 1) * 8, py = ty * 16 + (lane >> 1);
    bool sky = true;
    const bool vec = (((uintptr_t)viewZ.data | ((size_t)viewZ.pitch * 4u)) & 15u) == 0 && px0 + 8 <= viewZ.w && py < viewZ.h;
    if (vec) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(viewZ.ptr<float>(px0, py))), b = __ldg(reinterpret_cast<const float4*>(viewZ.ptr<float>(px0 + 4, py)));
        sky = isSky(a.x) && isSky(a.y) && isSky(a.z) && isSky(a.w) && isSky(b.x) && isSky(b.y) && isSky(b.z) && isSky(b.w);
    } else {
#pragma unroll
        for (int i = 0; i < 8; i++) sky = sky && isSky(viewZ.load(px0 + i, py));
    }
    return __all_sync(0xFFFFFFFFu, sky);
}

struct TexR16F : TexView {
    NRD_DEV float fetch(int x, int y) const { return __half2float(__ushort_as_half(__ldg(ptr<unsigned short>(x, y)))); }
    NRD_DEV float load(int x, int y) const { return inside(x, y) ? fetch(x, y) : 0.0f; }
    NRD_DEV float fetchClamped(int x, int y) const { return fetch(cx(x), cy(y)); }
    NRD_DEV void store(int x, int y, float v) const { if (inside(x, y)) *ptrw<__half>(x, y) = __float2half_rn(v); }
    NRD_DEV float sampleNearest(float2 uv) const { return fetchClamped((int)floorf(uv.x * (float)w), (int)floorf(uv.y * (float)h)); }
    NRD_DEV float sampleLinear(float2 uv) const {
        float tx = uv.x * (float)w - 0.5f, ty = uv.y * (float)h - 0.5f;
        float fx = floorf(tx), fy = floorf(ty);
        float wx = tx - fx, wy = ty - fy;
        int x0 = (int)fx, y0 = (int)fy;
        float a = fetchClamped(x0, y0), b = fetchClamped(x0 + 1, y0), c = fetchClamped(x0, y0 + 1), d = fetchClamped(x0 + 1, y0 + 1);
        return lerp(lerp(a, b, wx), lerp(c, d, wx), wy);
    }
};

struct TexRGBA16F : TexView {
    NRD_DEV uint2 fetchRaw(int x, int y) const { return __ldg(ptr<uint2>(x, y)); }
    static NRD_DEV float4 decode(uint2 raw) {
        float2 lo = __half22float2(*reinterpret_cast<__half2*>(&raw.x));
        float2 hi = __half22float2(*reinterpret_cast<__half2*>(&raw.y));
        return make_float4(lo.x, lo.y, hi.x, hi.y);
    }
    NRD_DEV float4 fetch(int x, int y) const { return decode(fetchRaw(x, y)); }
    NRD_DEV float4 load(int x, int y) const { return inside(x, y) ? fetch(x, y) : f4(0.0f); }
    NRD_DEV float4 fetchClamped(int x, int y) const { return fetch(cx(x), cy(y)); }
    NRD_DEV float4 sampleNearest(float2 uv) const { return fetchClamped((int)floorf(uv.x * (float)w), (int)floorf(uv.y * (float)h)); }
    NRD_DEV void store(int x, int y, float4 v) const {
        if (!inside(x, y)) return;
        __half2 lo = __floats2half2_rn(v.x, v.y), hi = __floats2half2_rn(v.z, v.w);
        uint2 raw;
        raw.x = *reinterpret_cast<uint32_t*>(&lo);
        raw.y = *reinterpret_cast<uint32_t*>(&hi);
        *ptrw<uint2>(x, y) = raw;
    }
    NRD_DEV float4 sampleLinear(float2 uv) const {
        float tx = uv.x * (float)w - 0.5f, ty = uv.y * (float)h - 0.5f;
        float fx = floorf(tx), fy = floorf(ty);
        float wx = tx - fx, wy = ty - fy;
        int x0 = (int)fx, y0 = (int)fy;
        float4 a = fetchClamped(x0, y0), b = fetchClamped(x0 + 1, y0), c = fetchClamped(x0, y0 + 1), d = fetchClamped(x0 + 1, y0 + 1);
        return lerp(lerp(a, b, wx), lerp(c, d, wx), wy);
    }
};

// R10G10B10A2_UNORM normal + roughness + material
struct TexNR : TexView {
    NRD_DEV uint32_t fetchRaw(int x, int y) const { return __ldg(ptr<uint32_t>(x, y)); }
    NRD_DEV uint32_t loadRaw(int x, int y) const { return inside(x, y) ? fetchRaw(x, y) : 0u; }
    NRD_DEV uint32_t fetchRawClamped(int x, int y) const { return fetchRaw(cx(x), cy(y)); }
    NRD_DEV uint32_t sampleNearestRaw(float2 uv) const { return fetchRawClamped((int)floorf(uv.x * (float)w), (int)floorf(uv.y * (float)h)); }
    NRD_DEV void storeRaw(int x, int y, uint32_t v) const { if (inside(x, y)) *ptrw<uint32_t>(x, y) = v; }
    static NRD_DEV float4 decode(uint32_t v) {
        return make_float4((float)(v & 1023u) / 1023.0f, (float)((v >> 10) & 1023u) / 1023.0f, (float)((v >> 20) & 1023u) / 1023.0f, (float)(v >> 30) / 3.0f);
    }
};

// Executor-private geometry plane ( RGBA32F ): { world-space normal as unpackNormalRoughness returns it, | viewZ * viewZScale | } per texel of the
// current frame's IN_NORMAL_ROUGHNESS / IN_VIEWZ, written once per frame by reblurGeometryPlaneKernel and gathered by the spatial passes
struct TexGeom : TexView {
    NRD_DEV float4 fetch(int x, int y) const { return __ldg(ptr<float4>(x, y)); }
    NRD_DEV void store(int x, int y, float4 v) const { if (inside(x, y)) *ptrw<float4>(x, y) = v; }
};

// The sky-tile mask ( 0 or 1 per 16x16 tile ). R8_UNORM everywhere except REBLUR_DIFFUSE_SPECULAR_SH, whose pool table hands the passes a
// full-resolution RGBA16F texture under the TILES index ( Reblur_DiffuseSpecularSh.hpp:59-82: eleven textures for ten names ); graphics
// hardware converts on access, here the view does: `texelBytes` = 1 or 8, the mask lives in .x ( readers only ask "!= 0" ).
struct TexTiles : TexView {
    uint32_t texelBytes;
    NRD_DEV float load(int x, int y) const {
        if (!inside(x, y)) return 0.0f;
        const uint8_t* at = data + (size_t)(y * pitch + x) * texelBytes;
        return texelBytes == 1u ? (float)__ldg(at) : (float)__ldg(reinterpret_cast<const unsigned short*>(at));  // half bits of .x: zero iff the value is +0
    }
    NRD_DEV void store(int x, int y, float v) const {
        if (!inside(x, y)) return;
        uint8_t* at = data + (size_t)(y * pitch + x) * texelBytes;
        if (texelBytes == 1u) *at = (uint8_t)unormQ(v, 255.0f);
        else *reinterpret_cast<uint2*>(at) = make_uint2((uint32_t)__half_as_ushort(__float2half_rn(v)), 0u);  // a scalar store to a 4-channel UAV: ( v, 0, 0, 0 )
    }
};

struct TexR8 : TexView {  // R8_UNORM
    NRD_DEV float fetch(int x, int y) const { return (float)__ldg(ptr<uint8_t>(x, y)) / 255.0f; }
    NRD_DEV float load(int x, int y) const { return inside(x, y) ? fetch(x, y) : 0.0f; }
    NRD_DEV float fetchClamped(int x, int y) const { return fetch(cx(x), cy(y)); }
    NRD_DEV void store(int x, int y, float v) const { if (inside(x, y)) *ptrw<uint8_t>(x, y) = (uint8_t)unormQ(v, 255.0f); }
    NRD_DEV float sampleNearest(float2 uv) const { return fetchClamped((int)floorf(uv.x * (float)w), (int)floorf(uv.y * (float)h)); }
    NRD_DEV float sampleLinear(float2 uv) const {
        float tx = uv.x * (float)w - 0.5f, ty = uv.y * (float)h - 0.5f;
        float fx = floorf(tx), fy = floorf(ty);
        float wx = tx - fx, wy = ty - fy;
        int x0 = (int)fx, y0 = (int)fy;
        float a = fetchClamped(x0, y0), b = fetchClamped(x0 + 1, y0), c = fetchClamped(x0, y0 + 1), d = fetchClamped(x0 + 1, y0 + 1);
        return lerp(lerp(a, b, wx), lerp(c, d, wx), wy);
    }
};

struct TexRG8 : TexView {  // RG8_UNORM
    NRD_DEV float2 load(int x, int y) const {
        if (!inside(x, y)) return f2(0.0f);
        uchar2 v = __ldg(ptr<uchar2>(x, y));
        return make_float2((float)v.x / 255.0f, (float)v.y / 255.0f);
    }
    NRD_DEV void store(int x, int y, float2 v) const {
        if (inside(x, y)) *ptrw<uchar2>(x, y) = make_uchar2((unsigned char)unormQ(v.x, 255.0f), (unsigned char)unormQ(v.y, 255.0f));
    }
    NRD_DEV float2 fetchClamped(int x, int y) const {
        uchar2 v = __ldg(ptr<uchar2>(cx(x), cy(y)));
        return make_float2((float)v.x / 255.0f, (float)v.y / 255.0f);
    }
    NRD_DEV float2 sampleLinear(float2 uv) const {
        float tx = uv.x * (float)w - 0.5f, ty = uv.y * (float)h - 0.5f;
        float fx = floorf(tx), fy = floorf(ty);
        float wx = tx - fx, wy = ty - fy;
        int x0 = (int)fx, y0 = (int)fy;
        float2 a = fetchClamped(x0, y0), b = fetchClamped(x0 + 1, y0), c = fetchClamped(x0, y0 + 1), d = fetchClamped(x0 + 1, y0 + 1);
        return lerp(lerp(a, b, wx), lerp(c, d, wx), wy);
    }
};

struct TexRGBA8 : TexView {  // RGBA8_UNORM
    NRD_DEV float4 fetch(int x, int y) const {
        uchar4 v = __ldg(ptr<uchar4>(x, y));
        return make_float4((float)v.x / 255.0f, (float)v.y / 255.0f, (float)v.z / 255.0f, (float)v.w / 255.0f);
    }
    NRD_DEV float4 load(int x, int y) const { return inside(x, y) ? fetch(x, y) : f4(0.0f); }
    NRD_DEV float4 fetchClamped(int x, int y) const { return fetch(cx(x), cy(y)); }
    NRD_DEV float4 sampleNearest(float2 uv) const { return fetchClamped((int)floorf(uv.x * (float)w), (int)floorf(uv.y * (float)h)); }
    NRD_DEV float4 sampleLinear(float2 uv) const {
        float tx = uv.x * (float)w - 0.5f, ty = uv.y * (float)h - 0.5f;
        float fx = floorf(tx), fy = floorf(ty);
        float wx = tx - fx, wy = ty - fy;
        int x0 = (int)fx, y0 = (int)fy;
        float4 a = fetchClamped(x0, y0), b = fetchClamped(x0 + 1, y0), c = fetchClamped(x0, y0 + 1), d = fetchClamped(x0 + 1, y0 + 1);
        return lerp(lerp(a, b, wx), lerp(c, d, wx), wy);
    }
    NRD_DEV void store(int x, int y, float4 v) const {
        if (inside(x, y))
            *ptrw<uchar4>(x, y) = make_uchar4((unsigned char)unormQ(v.x, 255.0f), (unsigned char)unormQ(v.y, 255.0f), (unsigned char)unormQ(v.z, 255.0f),
                                              (unsigned char)unormQ(v.w, 255.0f));
    }
};

// First channel of an application-owned guide texture (IN_DIFF_CONFIDENCE / IN_SPEC_CONFIDENCE / IN_DISOCCLUSION_THRESHOLD_MIX): the
// reference declares them Texture2D<float>, so the application may bind any format and any size (NRDSample binds an RGBA16F texture at
// SHARC resolution). Read a few times per pixel at most, so the format switch is a uniform branch, not a template parameter.
struct TexAnyX : TexView {
    enum Kind : uint32_t { UNORM8 = 0, UNORM16 = 1, HALF = 2, FLOAT = 3 };
    uint32_t kind, bytesPerTexel;  // `pitch` stays in texels
    NRD_DEV float fetch(int x, int y) const {
        const uint8_t* at = data + (size_t)(y * pitch + x) * bytesPerTexel;
        switch (kind) {
            case UNORM8: return (float)__ldg(at) / 255.0f;
            case UNORM16: return (float)__ldg(reinterpret_cast<const unsigned short*>(at)) / 65535.0f;
            case HALF: return __half2float(__ushort_as_half(__ldg(reinterpret_cast<const unsigned short*>(at))));
            default: return __ldg(reinterpret_cast<const float*>(at));
        }
    }
    NRD_DEV float load(int x, int y) const { return inside(x, y) ? fetch(x, y) : 0.0f; }
    NRD_DEV float fetchClamped(int x, int y) const { return fetch(cx(x), cy(y)); }
    NRD_DEV float sampleLinear(float2 uv) const {
        float tx = uv.x * (float)w - 0.5f, ty = uv.y * (float)h - 0.5f;
        float fx = floorf(tx), fy = floorf(ty);
        float wx = tx - fx, wy = ty - fy;
        int x0 = (int)fx, y0 = (int)fy;
        float a = fetchClamped(x0, y0), b = fetchClamped(x0 + 1, y0), c = fetchClamped(x0, y0 + 1), d = fetchClamped(x0 + 1, y0 + 1);
        return lerp(lerp(a, b, wx), lerp(c, d, wx), wy);
    }
};

// Host side: view of an application texture as a single-channel guide; false if the format is not one of the supported ones
inline bool bindGuide(uint32_t format, void* data, uint32_t width, uint32_t height, uint32_t pitchBytes, TexAnyX& v) {
    using F = nrd::Format;
    uint32_t kind, bpp;
    switch ((F)format) {
        case F::R8_UNORM: kind = TexAnyX::UNORM8; bpp = 1; break;
        case F::RG8_UNORM: kind = TexAnyX::UNORM8; bpp = 2; break;
        case F::RGBA8_UNORM: kind = TexAnyX::UNORM8; bpp = 4; break;
        case F::R16_UNORM: kind = TexAnyX::UNORM16; bpp = 2; break;
        case F::RG16_UNORM: kind = TexAnyX::UNORM16; bpp = 4; break;
        case F::RGBA16_UNORM: kind = TexAnyX::UNORM16; bpp = 8; break;
        case F::R16_SFLOAT: kind = TexAnyX::HALF; bpp = 2; break;
        case F::RG16_SFLOAT: kind = TexAnyX::HALF; bpp = 4; break;
        case F::RGBA16_SFLOAT: kind = TexAnyX::HALF; bpp = 8; break;
        case F::R32_SFLOAT: kind = TexAnyX::FLOAT; bpp = 4; break;
        case F::RG32_SFLOAT: kind = TexAnyX::FLOAT; bpp = 8; break;
        case F::RGBA32_SFLOAT: kind = TexAnyX::FLOAT; bpp = 16; break;
        default: return false;
    }
    if (!data || (pitchBytes % bpp) != 0 || pitchBytes < width * bpp) return false;
    v.data = (uint8_t*)data;
    v.w = (int)width;
    v.h = (int)height;
    v.pitch = (int)(pitchBytes / bpp);
    v.kind = kind;
    v.bytesPerTexel = bpp;
    return true;
}

struct TexR16U : TexView {
    NRD_DEV uint32_t fetch(int x, int y) const { return __ldg(ptr<unsigned short>(x, y)); }
    NRD_DEV uint32_t fetchClamped(int x, int y) const { return fetch(cx(x), cy(y)); }
    NRD_DEV void store(int x, int y, uint32_t v) const { if (inside(x, y)) *ptrw<unsigned short>(x, y) = (unsigned short)v; }
};

struct TexR8U : TexView {  // R8_UINT; integer stores saturate to the format's range like D3D / Vulkan image stores
    NRD_DEV uint32_t load(int x, int y) const { return inside(x, y) ? (uint32_t)__ldg(ptr<uint8_t>(x, y)) : 0u; }
    NRD_DEV void store(int x, int y, uint32_t v) const { if (inside(x, y)) *ptrw<uint8_t>(x, y) = (uint8_t)(v > 255u ? 255u : v); }
};

struct TexR32U : TexView {
    NRD_DEV uint32_t load(int x, int y) const { return inside(x, y) ? __ldg(ptr<uint32_t>(x, y)) : 0u; }
    NRD_DEV uint32_t fetchClamped(int x, int y) const { return __ldg(ptr<uint32_t>(cx(x), cy(y))); }
    NRD_DEV void store(int x, int y, uint32_t v) const { if (inside(x, y)) *ptrw<uint32_t>(x, y) = v; }
};

// =================================================================================================================
// Matrices as they sit in the constant buffer (column-major)
// =================================================================================================================
NRD_DEV float4 mulM4(const Mat4& M, float4 v) {
    return make_float4(M.m[0] * v.x + M.m[4] * v.y + M.m[8] * v.z + M.m[12] * v.w, M.m[1] * v.x + M.m[5] * v.y + M.m[9] * v.z + M.m[13] * v.w,
                       M.m[2] * v.x + M.m[6] * v.y + M.m[10] * v.z + M.m[14] * v.w, M.m[3] * v.x + M.m[7] * v.y + M.m[11] * v.z + M.m[15] * v.w);
}
NRD_DEV float3 rotate(const Mat4& M, float3 v) {  // (float3x3)M * v
    return make_float3(M.m[0] * v.x + M.m[4] * v.y + M.m[8] * v.z, M.m[1] * v.x + M.m[5] * v.y + M.m[9] * v.z, M.m[2] * v.x + M.m[6] * v.y + M.m[10] * v.z);
}
NRD_DEV float3 rotateInverse(const Mat4& M, float3 v) {  // transpose((float3x3)M) * v
    return make_float3(M.m[0] * v.x + M.m[1] * v.y + M.m[2] * v.z, M.m[4] * v.x + M.m[5] * v.y + M.m[6] * v.z, M.m[8] * v.x + M.m[9] * v.y + M.m[10] * v.z);
}
NRD_DEV float3 affine(const Mat4& M, float3 p) { return xyz(mulM4(M, f4(p, 1.0f))); }
NRD_DEV float2 screenUv(const Mat4& worldToClip, float3 X) {  // Geometry::GetScreenUv, D3D origin
    float4 clip = mulM4(worldToClip, f4(X, 1.0f));
    return make_float2(clip.x / clip.w, clip.y / clip.w) * make_float2(0.5f, -0.5f) + 0.5f;
}

// =================================================================================================================
// Math
// =================================================================================================================
NRD_DEV float linearStep(float a, float b, float x) { return saturate((x - a) / (b - a)); }
// ( 3 - 2 x as one FFMA: x * 2 is exact, so the bits are those of 3.0f - x * 2.0f; nvcc otherwise emits x + x and a subtraction )
NRD_DEV float smoothStep01(float x) { x = saturate(x); return (x * x) * fmaf(x, -2.0f, 3.0f); }
NRD_DEV float smoothStep(float a, float b, float x) { x = linearStep(a, b, x); return (x * x) * fmaf(x, -2.0f, 3.0f); }
NRD_DEV float pow01(float x, float y) { return powf(saturate(x), y); }
NRD_DEV float sqrt01(float x) { return sqrtf(saturate(x)); }
NRD_DEV float rsqrtSafe(float x) { return 1.0f / sqrtf(fmaxf(x, ML_SMALL_EPS)); }
NRD_DEV float positiveRcp(float x) { return 1.0f / fmaxf(x, ML_SMALL_EPS); }
NRD_DEV float acosApproxPositive(float x) { return lerp(1.567589f, 1.399331f, saturate(x)) * sqrtf(saturate(1.0f - x)); }
NRD_DEV float degToRad(float x) { return x * (3.14159265358979323846f / 180.0f); }

NRD_DEV float4 scaleRotator(float4 r, float2 s) { return make_float4(s.x * r.x, s.x * r.z, s.y * r.y, s.y * r.w); }
NRD_DEV float2 rotate2(float4 r, float2 v) { return make_float2(v.x * r.x + v.y * r.y, v.x * r.z + v.y * r.w); }

struct Basis { float3 T, B, N; };
NRD_DEV Basis getBasis(float3 N) {
    float sz = signFast(N.z);
    float a = 1.0f / (sz + N.z);
    float ya = N.y * a;
    float b = N.x * ya;
    float c = N.x * sz;
    Basis r;
    r.T = make_float3(c * N.x * a - 1.0f, sz * b, c);
    r.B = make_float3(b, N.y * ya - sz, N.y);
    r.N = N;
    return r;
}

NRD_DEV float3 reconstructViewPosition(float2 uv, const float* frustum, float viewZ, float orthoMode) {
    float s = orthoMode == 0.0f ? viewZ : orthoMode;
    return make_float3((uv.x * frustum[2] + frustum[0]) * s, (uv.y * frustum[3] + frustum[1]) * s, viewZ);
}

// Packing::RgbaToUint( c, 6, 6, 4, 0 ) / UintToRgba for the 16-bit "internal data" word
NRD_DEV uint32_t packInternal664(float r, float g, float b) {
    return (uint32_t)(saturate(r) * 63.0f + 0.5f) | ((uint32_t)(saturate(g) * 63.0f + 0.5f) << 6) | ((uint32_t)(saturate(b) * 15.0f + 0.5f) << 12);
}
NRD_DEV float3 unpackInternalData(uint32_t p) {  // REBLUR_Common.hlsli:29-38 -> ( diff frames, spec frames, materialID )
    float3 t = make_float3((float)(p & 63u) * (1.0f / 63.0f), (float)((p >> 6) & 63u) * (1.0f / 63.0f), (float)((p >> 12) & 15u) * (1.0f / 15.0f));
    return make_float3(roundNe(t.x * 63.0f), roundNe(t.y * 63.0f), t.z * 15.0f);
}

struct Bilinear { float2 origin, weights; };
NRD_DEV Bilinear getBilinearFilter(float2 uv, float2 texSize) {
    float2 t = uv * texSize - 0.5f;
    Bilinear r;
    r.origin = floor2(t);
    r.weights = saturate(t - r.origin);
    return r;
}
NRD_DEV float applyBilinear(float s00, float s10, float s01, float s11, Bilinear f) {
    return lerp(lerp(s00, s10, f.weights.x), lerp(s01, s11, f.weights.x), f.weights.y);
}
NRD_DEV float4 applyBilinear(float4 s00, float4 s10, float4 s01, float4 s11, Bilinear f) {
    return lerp(lerp(s00, s10, f.weights.x), lerp(s01, s11, f.weights.x), f.weights.y);
}
NRD_DEV float4 bilinearCustomWeights(Bilinear f, float4 custom) {
    float2 o = saturate(1.0f - f.weights);
    return make_float4(custom.x * (o.x * o.y), custom.y * (f.weights.x * o.y), custom.z * (o.x * f.weights.y), custom.w * (f.weights.x * f.weights.y));
}
NRD_DEV float applyCustomWeights(float s00, float s10, float s01, float s11, float4 w) {
    float sum = sum4(w);
    return (s00 * w.x + s10 * w.y + s01 * w.z + s11 * w.w) * (sum < 0.0001f ? 0.0f : 1.0f / sum);
}
NRD_DEV float modifiedRoughnessFromNormalVariance(float roughness, float3 avgNormal) {
    float l = length(avgNormal);
    float kappa = saturate(1.0f - l * l) * positiveRcp(l * (3.0f - l * l));
    return sqrt01(roughness * roughness + kappa);
}

// Sequence::Hash / HashCombine / Zorder and Rng::Hash
NRD_DEV uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
    return x;
}
NRD_DEV uint32_t explodeBits(uint32_t x) {
    x = (x | (x << 8)) & 0x00FF00FFu; x = (x | (x << 4)) & 0x0F0F0F0Fu; x = (x | (x << 2)) & 0x33333333u; x = (x | (x << 1)) & 0x55555555u;
    return x;
}
struct Rng {
    uint32_t state;
    NRD_DEV void init(uint32_t x, uint32_t y, uint32_t frame) {
        uint32_t seed = hash32(frame + 0x035F9F29u);
        uint32_t value = explodeBits(x) | (explodeBits(y) << 1);
        state = seed ^ (hash32(value) + 0x9E3779B9u + (seed << 6) + (seed >> 2));
    }
    NRD_DEV float next() {
        state = hash32(state);
        return __uint_as_float((state >> 9) | 0x3F800000u) - 1.0f;
    }
};

NRD_DEV float specularLobeTanHalfAngle(float roughness, float percentOfVolume) {
    percentOfVolume = saturate(percentOfVolume);
    return saturate(roughness) * sqrtf(percentOfVolume / (1.0f - percentOfVolume + ML_EPS));
}
NRD_DEV float4 specularDominantDirectionG2(float3 N, float3 V, float roughness) {
    float NoV = fabsf(dot(N, V));
    roughness = saturate(roughness);
    float a = 0.298475f * logf(39.4115f - 39.0029f * roughness);
    float factor = saturate(pow01(1.0f - NoV, 10.8649f) * (1.0f - a) + a);
    float3 R = reflect(-V, N);
    return f4(normalize(lerp(N, R, factor)), factor);
}

// NRD.hlsli
// materialID = alpha * 3 with alpha the 2-bit UNORM channel; ONE expression everywhere so equal IDs compare equal
NRD_DEV float materialFromRaw(uint32_t raw) { return (float)(raw >> 30) * (1.0f / 3.0f) * 3.0f; }
NRD_DEV float4 unpackNormalRoughness(uint32_t raw, float& materialID) {
    float4 p = TexNR::decode(raw);
    float t = p.z * 2.0f - 1.0f;
    float3 n;
    n.x = p.x - p.y;
    n.y = p.x + p.y - 1.0f;
    n.z = (t < 0.0f ? -1.0f : 1.0f) * (1.0f - fabsf(n.x) - fabsf(n.y));
    materialID = materialFromRaw(raw);
    n = n * (1.0f / sqrtf(dot(n, n) + 1e-9f));
    return f4(n, fabsf(t));
}
NRD_DEV float4 unpackNormalRoughness(uint32_t raw) { float m; return unpackNormalRoughness(raw, m); }
NRD_DEV float roughnessFromRaw(uint32_t raw) { return fabsf((float)((raw >> 20) & 1023u) / 1023.0f * 2.0f - 1.0f); }
NRD_DEV float3 linearToYCoCg(float3 c) {
    return make_float3(dot(c, make_float3(0.25f, 0.5f, 0.25f)), dot(c, make_float3(0.5f, 0.0f, -0.5f)), dot(c, make_float3(-0.25f, 0.5f, -0.25f)));
}
NRD_DEV float3 yCoCgToLinear(float3 c) {
    float t = c.x - c.z;
    return max3(make_float3(t + c.y, c.x + c.z, t - c.y), f3(0.0f));
}
NRD_DEV float specMagicCurve(float roughness, float power = 0.25f) {
    return (1.0f - exp2f(-200.0f * roughness * roughness)) * powf(saturate(roughness), power);
}
NRD_DEV float hitDistanceNormalization(float viewZ, const float* p, float roughness) {
    float smc = specMagicCurve(roughness, 0.5f);
    return (p[0] + fabsf(viewZ) * p[1]) * lerp(p[2], 1.0f, smc);
}

// Common.hlsli
NRD_DEV float stdDev(float m1, float m2) { return sqrtf(fabsf(m2 - m1 * m1)); }
NRD_DEV bool compareMaterials(float m0, float m, float minm) { return fmaxf(m0, minm) == fmaxf(m, minm); }
NRD_DEV float pixelRadiusToWorld(float unproject, float orthoMode, float pixelRadius, float viewZ) { return pixelRadius * unproject * lerp(viewZ, 1.0f, fabsf(orthoMode)); }
NRD_DEV float frustumSizeAt(float minRectDimMulUnproject, float orthoMode, float viewZ) { return minRectDimMulUnproject * lerp(viewZ, 1.0f, fabsf(orthoMode)); }
NRD_DEV float hitDistFactor(float hitDist, float frustumSize) { return saturate(hitDist / frustumSize); }
NRD_DEV bool isInScreenNearest(float2 uv) { return uv.x > 0.0f && uv.y > 0.0f && uv.x < 1.0f && uv.y < 1.0f; }
NRD_DEV float2 mirrorUv(float2 uv) {
    float2 m = 1.0f - fabs2(1.0f - frac2(uv * 0.5f) * 2.0f);
    return min2(m, f2(0.99999f));
}
NRD_DEV float4 isInScreenBilinear(float2 origin, float2 rectSize) {
    float x0 = (origin.x >= 0.0f && origin.x < rectSize.x) ? 1.0f : 0.0f, y0 = (origin.y >= 0.0f && origin.y < rectSize.y) ? 1.0f : 0.0f;
    float x1 = (origin.x + 1.0f >= 0.0f && origin.x + 1.0f < rectSize.x) ? 1.0f : 0.0f, y1 = (origin.y + 1.0f >= 0.0f && origin.y + 1.0f < rectSize.y) ? 1.0f : 0.0f;
    return make_float4(x0 * y0, x1 * y0, x0 * y1, x1 * y1);
}
NRD_DEV float parallaxInPixels(float3 X, float2 uvForZeroParallax, const Mat4& worldToClip, float2 rectSize) {
    return length((screenUv(worldToClip, X) - uvForZeroParallax) * rectSize);
}
NRD_DEV float3 getXvirtual(float hitDist, float curvature, float3 X, float3 Xprev, float3 N, float3 V, float roughness) {
    float4 D = specularDominantDirectionG2(N, V, roughness);
    float3 ray = xyz(D) * hitDist;
    Basis b = getBasis(N);
    float3 O = make_float3(dot(b.T, ray), dot(b.B, ray), -dot(b.N, ray));
    float mag = 1.0f / (2.0f * curvature * O.z - 1.0f);
    float NoV = fabsf(dot(N, V));
    float f = length(X);
    f *= saturate(1.0f - NoV);
    f *= fmaxf(curvature, 0.0f);
    f = 1.0f / (1.0f + f);
    mag *= f;
    float3 I = O * mag;
    float dw = D.w * length(I);
    float closeness = saturate(dw / (hitDist + NRD_EPS));
    float3 x = lerp(Xprev, X, closeness);
    return x + V * dw * signFast(mag);
}
NRD_DEV float normalWeightParam(float nonLinearAccumSpeed, float lobeAngleFraction, float roughness = 1.0f) {
    float percentOfVolume = NRD_MAX_PERCENT_OF_LOBE_VOLUME * lerp(saturate(lobeAngleFraction), 1.0f, nonLinearAccumSpeed);
    float tanHalfAngle = specularLobeTanHalfAngle(roughness, percentOfVolume);
    float angle = fmaxf(atanf(tanHalfAngle), NRD_NORMAL_ENCODING_ERROR);
    return 1.0f / angle;
}
NRD_DEV float2 geometryWeightParams(float planeDistSensitivity, float frustumSize, float3 Xv, float3 Nv) {
    float a = 1.0f / (planeDistSensitivity * frustumSize);
    return make_float2(a, -(dot(Nv, Xv) * a));
}
NRD_DEV float2 hitDistanceWeightParams(float hitDist, float nonLinearAccumSpeed) {
    float a = 1.0f / nonLinearAccumSpeed;
    return make_float2(a, -(hitDist * a));
}
NRD_DEV float2 roughnessWeightParams(float roughness, float fraction, float sensitivity = NRD_ROUGHNESS_SENSITIVITY) {
    float a = 1.0f / lerp(sensitivity, 1.0f, saturate(roughness * fraction));
    return make_float2(a, -(roughness * a));
}
NRD_DEV float2 relaxedRoughnessWeightParams(float m, float fraction = 1.0f, float sensitivity = NRD_ROUGHNESS_SENSITIVITY) {
    float a = 1.0f / lerp(sensitivity, 1.0f, lerp(m * m, m, saturate(fraction)));
    return make_float2(a, -(m * a));
}
NRD_DEV float expApprox(float x) { return 1.0f / (x * x - x + 1.0f); }
NRD_DEV float exponentialWeight(float x, float px, float py) { return expApprox(-3.0f * fabsf(x * px + py)); }
// Math::SmoothStep( 1, 0, |x px + py| ): ( t - 1 ) / ( 0 - 1 ) is exactly 1 - t, so the division is dropped
#ifdef NRD_B200_PLAIN_HELPERS   // experiments ( tools/build_variant.py ): the C forms nvcc turns into three + four instructions
NRD_DEV float nonExponentialWeight(float x, float px, float py) { float s = saturate(1.0f - fabsf(x * px + py)); return s * s * (3.0f - s * 2.0f); }
#else
NRD_DEV float nonExponentialWeight(float x, float px, float py) { float s = satOneMinusAbs(x * px + py); return (s * s) * fmaf(s, -2.0f, 3.0f); }
#endif
NRD_DEV float gaussianWeight(float r) { return expf(-0.66f * r * r); }
NRD_DEV float encodingAwareNormalWeight(float3 Ncurr, float3 Nprev, float maxAngle, float curvatureAngle, float thresholdAngle) {
    float angle = acosApproxPositive(dot(Ncurr, Nprev));
    float w = smoothStep01(1.0f - (angle - curvatureAngle - thresholdAngle) / maxAngle);
    return smoothStep(0.05f, 0.95f, w);
}
NRD_DEV float disocclusionThresholdAt(float threshold, float frustumSize, float NoV) { return frustumSize * saturate(threshold / fmaxf(0.05f, NoV)); }

}  // namespace nrdk
