// Front end / back end on the device: what a CUDA / OptiX renderer runs either side of the denoiser instead of the HLSL of NRD.hlsli
// (NRDSample: Shaders/TraceOpaque.cs.hlsl:738-757 packs, Shaders/Composition.cs.hlsl:85-118 unpacks). All arithmetic is include/nrd_frontend.cuh;
// these kernels only move data: one thread per pixel, fp32 float4 planes in (128-bit loads), API-format textures out, and back.
// HBM-bound by construction: 16-32 B read + 4-8 B written per pixel.
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <atomic>
#include <string>

#include "../../../include/nrd_frontend.cuh"
#include "../../../include/nrdcu.h"
#include "../frontend_probe.inl"

namespace {

using namespace nrdfe;

__global__ void __launch_bounds__(256) frontEndProbeKernel(const float4* a, const float4* b, const float4* c, const float4* d, const float4* e, const float4* f, float4* const* out,
                                                           int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 v[6] = {a[i], b[i], c[i], d[i], e[i], f[i]};
    F4 in[6], r[21];
    for (int k = 0; k < 6; k++) in[k] = f4(v[k].x, v[k].y, v[k].z, v[k].w);
    frontEndProbeColumn(in, r);
    for (int k = 0; k < 21; k++) out[k][i] = make_float4(r[k].x, r[k].y, r[k].z, r[k].w);
}

struct Plane { uint8_t* data; int w, h, pitch; };  // pitch in bytes
__device__ __forceinline__ uint2 toHalf4(F4 v) {
    __half2 lo = __floats2half2_rn(v.x, v.y), hi = __floats2half2_rn(v.z, v.w);
    return make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
}
__device__ __forceinline__ F4 fromHalf4(uint2 raw) {
    float2 lo = __half22float2(*reinterpret_cast<__half2*>(&raw.x)), hi = __half22float2(*reinterpret_cast<__half2*>(&raw.y));
    return f4(lo.x, lo.y, hi.x, hi.y);
}

// IN_NORMAL_ROUGHNESS (R10G10B10A2_UNORM) from { N.xyz, linear roughness } (+ optional material ID 0..3)
__global__ void __launch_bounds__(256) packNormalRoughnessKernel(const float4* __restrict__ normalRoughness, const float* __restrict__ materialID, Plane out) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= out.w) return;
    const size_t i = (size_t)y * out.w + x;
    const float4 v = normalRoughness[i];
    const F4 p = NRD_FrontEnd_PackNormalAndRoughness(f3(v.x, v.y, v.z), v.w, materialID ? materialID[i] : 0.0f);
    reinterpret_cast<uint32_t*>(out.data + (size_t)y * out.pitch)[x] = packR10G10B10A2(p);
}

// IN_DIFF / IN_SPEC_RADIANCE_HITDIST (RGBA16F) from { linear radiance, hit distance in world units }.
// mode 0: REBLUR ( YCoCg + hit distance normalised by REBLUR_FrontEnd_GetNormHitDist; roughness = 1 for the diffuse lobe, read from the packed G-buffer for specular )
// mode 1: RELAX ( radiance and hit distance as they are, sanitised )
__global__ void __launch_bounds__(256) packRadianceHitDistKernel(const float4* __restrict__ radianceHitDist, Plane viewZ, Plane normalRoughness, Plane out, int mode, int isSpecular,
                                                                 float hdA, float hdB, float hdC) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= out.w) return;
    const float4 v = radianceHitDist[(size_t)y * out.w + x];
    F4 p;
    if (mode == 0) {
        const float z = reinterpret_cast<const float*>(viewZ.data + (size_t)y * viewZ.pitch)[x];
        float roughness = 1.0f;
        if (isSpecular) roughness = NRD_FrontEnd_UnpackNormalAndRoughness(unpackR10G10B10A2(reinterpret_cast<const uint32_t*>(normalRoughness.data + (size_t)y * normalRoughness.pitch)[x])).w;
        const float normHitDist = REBLUR_FrontEnd_GetNormHitDist(v.w, z, f3(hdA, hdB, hdC), roughness);
        p = REBLUR_FrontEnd_PackRadianceAndNormHitDist(f3(v.x, v.y, v.z), normHitDist, true);
    } else
        p = RELAX_FrontEnd_PackRadianceAndHitDist(f3(v.x, v.y, v.z), v.w, true);
    reinterpret_cast<uint2*>(out.data + (size_t)y * out.pitch)[x] = toHalf4(p);
}

// OUT_DIFF / OUT_SPEC_RADIANCE_HITDIST (RGBA16F) -> { linear radiance, .w as stored }
__global__ void __launch_bounds__(256) unpackRadianceKernel(Plane in, float4* __restrict__ out, int mode) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= in.w) return;
    F4 v = fromHalf4(reinterpret_cast<const uint2*>(in.data + (size_t)y * in.pitch)[x]);
    v = mode == 0 ? REBLUR_BackEnd_UnpackRadianceAndNormHitDist(v) : RELAX_BackEnd_UnpackRadiance(v);
    out[(size_t)y * in.w + x] = make_float4(v.x, v.y, v.z, v.w);
}

thread_local std::string g_frontEndError;
uint32_t bad(const char* what) {
    g_frontEndError = what;
    return 2u;  // nrd::Result::INVALID_ARGUMENT
}
bool plane(const nrdcuTexture* t, uint32_t format, uint32_t bpp, Plane& p) {
    if (!t || !t->data || t->format != format || t->pitchBytes < t->width * bpp) return false;
    p = {(uint8_t*)t->data, (int)t->width, (int)t->height, (int)t->pitchBytes};
    return true;
}
uint32_t launched(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        g_frontEndError = std::string(what) + ": " + cudaGetErrorString(e);
        return 1u;  // FAILURE
    }
    return 0u;
}
}  // namespace

namespace nrdk { void countLaunch(); }

extern "C" {

NRDCU_API const char* nrdcuFrontEndGetLastError(void) { return g_frontEndError.c_str(); }

NRDCU_API uint32_t nrdcuFrontEndProbe(const float* const* in6, float* const* out21, uint32_t n, void* stream) {
    if (!in6 || !out21 || !n) return bad("nrdcuFrontEndProbe: null argument");
    float4** table = nullptr;
    if (cudaMalloc(&table, 21 * sizeof(float4*)) != cudaSuccess) return 1u;
    cudaMemcpyAsync(table, out21, 21 * sizeof(float4*), cudaMemcpyHostToDevice, (cudaStream_t)stream);
    frontEndProbeKernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>((const float4*)in6[0], (const float4*)in6[1], (const float4*)in6[2], (const float4*)in6[3],
                                                                         (const float4*)in6[4], (const float4*)in6[5], table, (int)n);
    uint32_t rc = launched("nrdcuFrontEndProbe");
    cudaStreamSynchronize((cudaStream_t)stream);
    cudaFree(table);
    if (!rc) nrdk::countLaunch();
    return rc;
}

NRDCU_API uint32_t nrdcuFrontEndPackNormalRoughness(const float* normalRoughness, const float* materialID, const nrdcuTexture* outNormalRoughness, void* stream) {
    Plane o;
    if (!normalRoughness || !plane(outNormalRoughness, 40u /* R10_G10_B10_A2_UNORM */, 4, o)) return bad("nrdcuFrontEndPackNormalRoughness: expects fp32 { N, roughness } and an R10_G10_B10_A2_UNORM texture");
    packNormalRoughnessKernel<<<dim3((o.w + 255) / 256, o.h), 256, 0, (cudaStream_t)stream>>>((const float4*)normalRoughness, materialID, o);
    uint32_t rc = launched("nrdcuFrontEndPackNormalRoughness");
    if (!rc) nrdk::countLaunch();
    return rc;
}

NRDCU_API uint32_t nrdcuFrontEndPackRadianceHitDist(uint32_t mode, const float* radianceHitDist, const nrdcuTexture* viewZ, const nrdcuTexture* normalRoughness, uint32_t isSpecular,
                                                    const float* hitDistParams3, const nrdcuTexture* out, void* stream) {
    Plane o, z = {}, nr = {};
    if (mode > 1u || !radianceHitDist || !plane(out, 27u /* RGBA16_SFLOAT */, 8, o)) return bad("nrdcuFrontEndPackRadianceHitDist: expects mode 0 ( REBLUR ) / 1 ( RELAX ), fp32 input and an RGBA16_SFLOAT texture");
    if (mode == 0u) {
        if (!hitDistParams3 || !plane(viewZ, 30u /* R32_SFLOAT */, 4, z) || z.w != o.w || z.h != o.h) return bad("nrdcuFrontEndPackRadianceHitDist: REBLUR needs IN_VIEWZ ( R32_SFLOAT ) and ReblurSettings::hitDistanceParameters");
        if (isSpecular && (!plane(normalRoughness, 40u, 4, nr) || nr.w != o.w || nr.h != o.h)) return bad("nrdcuFrontEndPackRadianceHitDist: the specular lobe needs IN_NORMAL_ROUGHNESS");
    }
    packRadianceHitDistKernel<<<dim3((o.w + 255) / 256, o.h), 256, 0, (cudaStream_t)stream>>>((const float4*)radianceHitDist, z, nr, o, (int)mode, isSpecular ? 1 : 0,
                                                                                              hitDistParams3 ? hitDistParams3[0] : 0.0f, hitDistParams3 ? hitDistParams3[1] : 0.0f,
                                                                                              hitDistParams3 ? hitDistParams3[2] : 0.0f);
    uint32_t rc = launched("nrdcuFrontEndPackRadianceHitDist");
    if (!rc) nrdk::countLaunch();
    return rc;
}

NRDCU_API uint32_t nrdcuBackEndUnpackRadiance(uint32_t mode, const nrdcuTexture* in, float* outRadiance, void* stream) {
    Plane i;
    if (mode > 1u || !outRadiance || !plane(in, 27u, 8, i)) return bad("nrdcuBackEndUnpackRadiance: expects mode 0 ( REBLUR ) / 1 ( RELAX ), an RGBA16_SFLOAT texture and an fp32 output");
    unpackRadianceKernel<<<dim3((i.w + 255) / 256, i.h), 256, 0, (cudaStream_t)stream>>>(i, (float4*)outRadiance, (int)mode);
    uint32_t rc = launched("nrdcuBackEndUnpackRadiance");
    if (!rc) nrdk::countLaunch();
    return rc;
}
}
