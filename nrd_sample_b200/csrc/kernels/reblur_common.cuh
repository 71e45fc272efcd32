// REBLUR helpers that read the per-frame constants (REBLUR_Common.hlsli:13-371, Common.hlsli:261-262, 567-572,
// 604-658) and the launch-parameter blocks of the REBLUR kernels.
#pragma once
#include <type_traits>

#include "common.cuh"
#include "pairmath.cuh"

namespace nrdk {

using nrdb::ReblurConstants;

constexpr float REBLUR_MAX_ACCUM_FRAME_NUM = 63.0f;
constexpr float REBLUR_MAX_MATERIALID_NUM = 15.0f;
constexpr float REBLUR_INVALID = -32768.0f;

// NRD_SIGNAL of a permutation (NRD.hlsli:338-339): kernels are templated on it, the bindings of the lobe a denoiser does not have
// (REBLUR_DIFFUSE / REBLUR_SPECULAR) stay zero-initialised and are never touched
enum : int { SIGNAL_DIFF = 1, SIGNAL_SPEC = 2, SIGNAL_BOTH = 3 };
// host side of the launchers: run `call( std::integral_constant< int, SIGNAL > )` for the run-time NRD_SIGNAL
template <class F> inline void withSignal(int signal, F&& call) {
    if (signal == SIGNAL_DIFF) call(std::integral_constant<int, SIGNAL_DIFF>());
    else if (signal == SIGNAL_SPEC) call(std::integral_constant<int, SIGNAL_SPEC>());
    else call(std::integral_constant<int, SIGNAL_BOTH>());
}

// NRD_MODE of a permutation ( NRD.hlsli ): RADIANCE and SH carry YCoCg radiance + hit distance in RGBA16F; OCCLUSION carries the normalized hit distance alone
// ( REBLUR_TYPE = float, REBLUR_Config.hlsli:103-107 ), DO a direction in .xyz + the hit distance in .w ( Reblur.cpp:36-39: R16_UNORM / RGBA16_SNORM pools,
// R8_UNORM fast history ). In registers every mode is a float4 with the hit distance in .w ( zeros in .xyz for OCCLUSION ).
enum : int { MODE_RADIANCE = 0, MODE_SH = 1, MODE_OCCLUSION = 2, MODE_DO = 3 };

// Format-polymorphic texel access for the occlusion modes: one uniform switch on the bound texture's nrd::Format ( pool textures are R16_UNORM /
// RGBA16_SNORM / R8_UNORM, application textures whatever the application owns ). RADIANCE / SH keep their fixed-width RGBA16F / R16F paths.
NRD_DEV float4 anyFetch4(const TexView& t, int x, int y) {
    const size_t i = (size_t)(y * t.pitch + x);
    switch ((nrd::Format)t.fmt) {
        case nrd::Format::R8_UNORM: return make_float4((float)__ldg(t.data + i) / 255.0f, 0.0f, 0.0f, 0.0f);
        case nrd::Format::R16_UNORM: return make_float4((float)__ldg(reinterpret_cast<const unsigned short*>(t.data) + i) / 65535.0f, 0.0f, 0.0f, 0.0f);
        case nrd::Format::R16_SFLOAT: return make_float4(__half2float(__ushort_as_half(__ldg(reinterpret_cast<const unsigned short*>(t.data) + i))), 0.0f, 0.0f, 0.0f);
        case nrd::Format::R32_SFLOAT: return make_float4(__ldg(reinterpret_cast<const float*>(t.data) + i), 0.0f, 0.0f, 0.0f);
        case nrd::Format::RGBA8_UNORM: {
            const uchar4 v = __ldg(reinterpret_cast<const uchar4*>(t.data) + i);
            return make_float4((float)v.x / 255.0f, (float)v.y / 255.0f, (float)v.z / 255.0f, (float)v.w / 255.0f);
        }
        case nrd::Format::RGBA16_UNORM: {
            const ushort4 v = __ldg(reinterpret_cast<const ushort4*>(t.data) + i);
            return make_float4((float)v.x / 65535.0f, (float)v.y / 65535.0f, (float)v.z / 65535.0f, (float)v.w / 65535.0f);
        }
        case nrd::Format::RGBA16_SNORM: {
            const short4 v = __ldg(reinterpret_cast<const short4*>(t.data) + i);
            return make_float4(fmaxf((float)v.x / 32767.0f, -1.0f), fmaxf((float)v.y / 32767.0f, -1.0f), fmaxf((float)v.z / 32767.0f, -1.0f), fmaxf((float)v.w / 32767.0f, -1.0f));
        }
        case nrd::Format::RGBA32_SFLOAT: return __ldg(reinterpret_cast<const float4*>(t.data) + i);
        default: return TexRGBA16F::decode(__ldg(reinterpret_cast<const uint2*>(t.data) + i));   // RGBA16_SFLOAT
    }
}
NRD_DEV int snormQ(float v) { v = fminf(fmaxf(v, -1.0f), 1.0f) * 32767.0f; return (int)(v + (v >= 0.0f ? 0.5f : -0.5f)); }   // NaN -> 0
NRD_DEV void anyStore4(const TexView& t, int x, int y, float4 v) {
    if (!t.inside(x, y)) return;
    const size_t i = (size_t)(y * t.pitch + x);
    switch ((nrd::Format)t.fmt) {
        case nrd::Format::R8_UNORM: t.data[i] = (uint8_t)unormQ(v.x, 255.0f); break;
        case nrd::Format::R16_UNORM: reinterpret_cast<unsigned short*>(t.data)[i] = (unsigned short)unormQ(v.x, 65535.0f); break;
        case nrd::Format::R16_SFLOAT: reinterpret_cast<__half*>(t.data)[i] = __float2half_rn(v.x); break;
        case nrd::Format::R32_SFLOAT: reinterpret_cast<float*>(t.data)[i] = v.x; break;
        case nrd::Format::RGBA8_UNORM:
            reinterpret_cast<uchar4*>(t.data)[i] = make_uchar4((unsigned char)unormQ(v.x, 255.0f), (unsigned char)unormQ(v.y, 255.0f), (unsigned char)unormQ(v.z, 255.0f), (unsigned char)unormQ(v.w, 255.0f));
            break;
        case nrd::Format::RGBA16_UNORM:
            reinterpret_cast<ushort4*>(t.data)[i] = make_ushort4((unsigned short)unormQ(v.x, 65535.0f), (unsigned short)unormQ(v.y, 65535.0f), (unsigned short)unormQ(v.z, 65535.0f), (unsigned short)unormQ(v.w, 65535.0f));
            break;
        case nrd::Format::RGBA16_SNORM: reinterpret_cast<short4*>(t.data)[i] = make_short4((short)snormQ(v.x), (short)snormQ(v.y), (short)snormQ(v.z), (short)snormQ(v.w)); break;
        case nrd::Format::RGBA32_SFLOAT: reinterpret_cast<float4*>(t.data)[i] = v; break;
        default: {
            const __half2 lo = __floats2half2_rn(v.x, v.y), hi = __floats2half2_rn(v.z, v.w);
            reinterpret_cast<uint2*>(t.data)[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&lo), *reinterpret_cast<const uint32_t*>(&hi));
        }
    }
}

// The lobe's signal texture ( bound through a TexRGBA16F-typed member in every mode ) and its fast history ( TexR16F-typed member )
template <int MODE> struct Sig {
    static constexpr bool FIXED = MODE == MODE_RADIANCE || MODE == MODE_SH;
    // OCCLUSION: Texture2D< float > reads .x of whatever is bound; in registers the value travels in .w
    static NRD_DEV float4 fromTexel(float4 v) { return MODE == MODE_OCCLUSION ? make_float4(0.0f, 0.0f, 0.0f, v.x) : v; }
    static NRD_DEV float4 toTexel(float4 v) { return MODE == MODE_OCCLUSION ? make_float4(v.w, 0.0f, 0.0f, 0.0f) : v; }
    static NRD_DEV float4 fetch(const TexRGBA16F& t, int x, int y) {
        if constexpr (FIXED) return t.fetch(x, y);
        else return fromTexel(anyFetch4(t, x, y));
    }
    static NRD_DEV float4 load(const TexRGBA16F& t, int x, int y) { return t.inside(x, y) ? fetch(t, x, y) : f4(0.0f); }
    static NRD_DEV float4 fetchClamped(const TexRGBA16F& t, int x, int y) { return fetch(t, t.cx(x), t.cy(y)); }
    static NRD_DEV void store(const TexRGBA16F& t, int x, int y, float4 v) {
        if constexpr (FIXED) t.store(x, y, v);
        else anyStore4(t, x, y, toTexel(v));
    }
    // REBLUR_Common.hlsli:147-238
    static NRD_DEV float luma(float4 c) { return FIXED ? c.x : c.w; }
};
template <int MODE> struct FastSig {   // REBLUR_FAST_TYPE: R16F luma, or the R8_UNORM hit distance of the occlusion modes
    static NRD_DEV float fetch(const TexR16F& t, int x, int y) {
        if constexpr (Sig<MODE>::FIXED) return t.fetch(x, y);
        else return anyFetch4(t, x, y).x;
    }
    static NRD_DEV float load(const TexR16F& t, int x, int y) { return t.inside(x, y) ? fetch(t, x, y) : 0.0f; }
    static NRD_DEV float fetchClamped(const TexR16F& t, int x, int y) { return fetch(t, t.cx(x), t.cy(y)); }
    static NRD_DEV void store(const TexR16F& t, int x, int y, float v) {
        if constexpr (Sig<MODE>::FIXED) t.store(x, y, v);
        else anyStore4(t, x, y, make_float4(v, 0.0f, 0.0f, 0.0f));
    }
};

NRD_DEV float unpackViewZ(const ReblurConstants& cb, float z) { return fabsf(z * cb.viewZScale); }
NRD_DEV bool inDenoisingRange(const ReblurConstants& cb, float z) { return z < cb.denoisingRange; }
NRD_DEV float applyGeometryWeightLast(const ReblurConstants& cb, float w, float z, float NoX, float2 p) {
    w *= nonExponentialWeight(NoX, p.x, p.y);
    return !inDenoisingRange(cb, z) ? 0.0f : w;
}

NRD_DEV uint32_t packInternalData(const ReblurConstants& cb, float diffAccumSpeed, float specAccumSpeed, float materialID) {
    diffAccumSpeed = fminf(diffAccumSpeed + 1.0f, cb.maxAccumulatedFrameNum);
    specAccumSpeed = fminf(specAccumSpeed + 1.0f, cb.maxAccumulatedFrameNum);
    return packInternal664(roundNe(diffAccumSpeed) / REBLUR_MAX_ACCUM_FRAME_NUM, roundNe(specAccumSpeed) / REBLUR_MAX_ACCUM_FRAME_NUM, materialID / REBLUR_MAX_MATERIALID_NUM);
}
NRD_DEV float2 packData1(float diffAccumSpeed, float specAccumSpeed) {
    return make_float2(saturate(roundNe(diffAccumSpeed) / REBLUR_MAX_ACCUM_FRAME_NUM), saturate(roundNe(specAccumSpeed) / REBLUR_MAX_ACCUM_FRAME_NUM));
}
NRD_DEV float2 unpackData1(float2 p) { return make_float2(roundNe(p.x * REBLUR_MAX_ACCUM_FRAME_NUM), roundNe(p.y * REBLUR_MAX_ACCUM_FRAME_NUM)); }
NRD_DEV uint32_t packData2(float fbits, float curvature, float virtualHistoryAmount, bool smbAllowCatRom) {
    uint32_t p = (uint32_t)(fbits + 0.5f);
    p |= (uint32_t)(saturate(virtualHistoryAmount) * 127.0f + 0.5f) << 8;
    p |= smbAllowCatRom ? (1u << 15) : 0u;
    p |= (uint32_t)__half_as_ushort(__float2half_rn(curvature)) << 16;
    return p;
}
NRD_DEV float2 unpackData2(uint32_t p, uint32_t& bits, bool& smbAllowCatRom) {
    bits = p & 0xFFu;
    smbAllowCatRom = (p & (1u << 15)) != 0;
    return make_float2((float)((p >> 8) & 127u) / 127.0f, __half2float(__ushort_as_half((unsigned short)(p >> 16))));
}
// Single-lobe denoisers keep data1 in an R8_UNORM texture holding their one accumulation speed ( REBLUR_Common.hlsli:48-51, 58-61 ) and a
// diffuse-only one keeps data2 in 8 bits with the CatRom flag in bit 4 instead of 15 ( :69-73 ). P = any launch-parameter block below.
template <int SIGNAL, class P> NRD_DEV float2 loadData1(const P& p, int x, int y) {
    if constexpr (SIGNAL == SIGNAL_BOTH) return unpackData1(p.data1.load(x, y));
    else {
        const float v = roundNe(p.data1R8.load(x, y) * REBLUR_MAX_ACCUM_FRAME_NUM);
        return make_float2(v, SIGNAL == SIGNAL_DIFF ? 0.0f : v);
    }
}
template <int SIGNAL, class P> NRD_DEV void storeData1(const P& p, int x, int y, float diffAccumSpeed, float specAccumSpeed) {
    const float2 r = packData1(diffAccumSpeed, specAccumSpeed);
    if constexpr (SIGNAL == SIGNAL_BOTH) p.outData1.store(x, y, r);
    else p.outData1R8.store(x, y, SIGNAL == SIGNAL_DIFF ? r.x : r.y);
}
template <int SIGNAL, class P> NRD_DEV void storeData2(const P& p, int x, int y, float fbits, float curvature, float virtualHistoryAmount, bool smbAllowCatRom) {
    if constexpr ((SIGNAL & SIGNAL_SPEC) != 0) p.outData2.store(x, y, packData2(fbits, curvature, virtualHistoryAmount, smbAllowCatRom));
    else p.outData2R8.store(x, y, (uint32_t)(fbits + 0.5f) | (smbAllowCatRom ? (1u << 4) : 0u));  // curvature = virtualHistoryAmount = 0
}
template <int SIGNAL, class P> NRD_DEV float2 loadData2(const P& p, int x, int y, uint32_t& bits, bool& smbAllowCatRom) {
    if constexpr ((SIGNAL & SIGNAL_SPEC) != 0) return unpackData2(p.data2.load(x, y), bits, smbAllowCatRom);
    else {
        bits = p.data2R8.load(x, y);
        smbAllowCatRom = (bits & (1u << 4)) != 0;
        return make_float2(0.0f, 0.0f);
    }
}
NRD_DEV float3 viewVector(const ReblurConstants& cb, float3 X, bool isViewSpace = false) {
    return cb.orthoMode == 0.0f ? normalize(-X) : (isViewSpace ? make_float3(0, 0, -1) : make_float3(cb.viewVectorWorld[0], cb.viewVectorWorld[1], cb.viewVectorWorld[2]));
}
NRD_DEV float3 viewVectorPrev(const ReblurConstants& cb, float3 Xprev, float3 cameraDelta) {
    return cb.orthoMode == 0.0f ? normalize(cameraDelta - Xprev) : make_float3(cb.viewVectorWorldPrev[0], cb.viewVectorWorldPrev[1], cb.viewVectorWorldPrev[2]);
}
NRD_DEV float minHitDistAccumSpeed(const ReblurConstants& cb, float roughness) {
    return 1.0f / (1.0f + 0.5f * specMagicCurve(roughness) * cb.maxAccumulatedFrameNum);
}
NRD_DEV float responsiveFactor(const ReblurConstants& cb, float roughness) { return smoothStep01(fmaxf(roughness, 1e-3f) * cb.responsiveAccumulationInvRoughnessThreshold); }
NRD_DEV float lumaScale(float currLuma, float newLuma) { return (newLuma + NRD_EPS) / (currLuma + NRD_EPS); }
NRD_DEV float4 mixHistoryAndCurrent(const ReblurConstants& cb, float4 history, float4 current, float f, float roughness = 1.0f) {
    float fw = fmaxf(f, minHitDistAccumSpeed(cb, roughness));
    return make_float4(lerp(history.x, current.x, f), lerp(history.y, current.y, f), lerp(history.z, current.z, f), lerp(history.w, current.w, fw));
}
// REBLUR_SH_TYPE sh; sh *= GetLumaScale( length( sh ), luma )
NRD_DEV float4 rescaleSh(float4 sh, float newLuma) { return sh * lumaScale(sqrtf(dot(sh, sh)), newLuma); }
NRD_DEV float4 changeLuma(float4 c, float newLuma) {
    float s = lumaScale(c.x, newLuma);
    return make_float4(c.x * s, c.y * s, c.z * s, c.w);
}
NRD_DEV float4 clampNegativeToZero(float4 c) { return f4(linearToYCoCg(yCoCgToLinear(xyz(c))), saturate(c.w)); }
// GetLuma / ChangeLuma / ClampNegativeToZero per NRD_MODE ( REBLUR_Common.hlsli:147-238 ): "luma" is the hit distance in the occlusion modes
template <int MODE> NRD_DEV float lumaOf(float4 c) { return Sig<MODE>::luma(c); }
template <int MODE> NRD_DEV float4 changeLumaM(float4 c, float newLuma) {
    if constexpr (MODE == MODE_OCCLUSION) return make_float4(0.0f, 0.0f, 0.0f, newLuma);
    else if constexpr (MODE == MODE_DO) {
        const float s = lumaScale(c.w, newLuma);
        return make_float4(c.x * s, c.y * s, c.z * s, newLuma);
    } else return changeLuma(c, newLuma);
}
template <int MODE> NRD_DEV float4 clampNegativeToZeroM(float4 c) {
    if constexpr (MODE == MODE_OCCLUSION) return make_float4(0.0f, 0.0f, 0.0f, saturate(c.w));
    else if constexpr (MODE == MODE_DO) return changeLumaM<MODE_DO>(c, saturate(c.w));
    else return clampNegativeToZero(c);
}
NRD_DEV float computeAntilag(const ReblurConstants& cb, float h, float a, float sigma, float accumSpeed) {
    float s = sigma * cb.antilagSettings[0];
    float magic = cb.antilagSettings[1] * cb.framerateScale * cb.framerateScale;
    float hc = clampf(h, a - s, a + s);
    float d = fabsf(h - hc) / (fmaxf(h, hc) + NRD_EPS);
    return 1.0f / (1.0f + d * accumSpeed / magic);
}
NRD_DEV void kernelBasis(float3 D, float3 N, float3& T, float3& B) {
    Basis basis = getBasis(N);
    T = basis.T;
    B = basis.B;
    if (fabsf(dot(D, N)) < 0.999f) {
        float3 R = reflect(-D, N);
        T = normalize(cross(N, R));
        B = cross(R, T);
    }
}
// pixels the checkerboard did not trace this frame accumulate faster ( REBLUR_Common.hlsli:302-303 )
NRD_DEV float checkerboardResolveAccumSpeed(const ReblurConstants& cb, float nonLinearAccumSpeed, bool hasData) {
    return hasData ? nonLinearAccumSpeed : nonLinearAccumSpeed * lerp(1.0f - cb.checkerboardResolveAccumSpeed, 1.0f, nonLinearAccumSpeed);
}
NRD_DEV float nonLinearAccumSpeedFast(const ReblurConstants& cb, float accumSpeed, float maxAccumSpeed, float confidence, bool hasData) {
    return checkerboardResolveAccumSpeed(cb, fmaxf(1.0f - confidence, 1.0f / (1.0f + fminf(accumSpeed, maxAccumSpeed))), hasData);
}
NRD_DEV float advancedNonLinearAccumSpeed(const ReblurConstants& cb, float accumSpeed) {
    float f = saturate(accumSpeed / (1.0f + cb.maxAccumulatedFrameNum * cb.convergenceSettings[2]));
    float e = cb.convergenceSettings[0] * lerp(cb.convergenceSettings[1], 1.0f, f);
    return 1.0f / (1.0f + e * accumSpeed);
}
NRD_DEV float2 temporalAccumulationParams(const ReblurConstants& cb, float footprintQuality, float accumSpeed, float antilag) {
    float w = footprintQuality;
    w *= 1.0f - advancedNonLinearAccumSpeed(cb, accumSpeed);
    w *= antilag;
    return make_float2(w, 1.0f + 3.0f * cb.framerateScale * w);
}

// 12-tap Catmull-Rom without corners, falling back to custom-weighted bilinear (Common.hlsli:604-658). The reference issues 5 hardware
// bilinear fetches at ( tc.x, -1 ), ( -1, tc.y ), ( tc.x, tc.y ), ( 2, tc.y ), ( tc.x, 2 ) around the footprint origin; four of them sit
// exactly on a texel row or column, so of the 20 texels a bilinear unit would touch only the 12 of the 4x4-minus-corners footprint carry
// weight. They are fetched directly ( clamp addressing, like the sampler ) and blended with the same lerps the sampler would do: 12
// loads + 8 lerps per texture instead of 20 loads + 15 lerps. The bilinear fallback reads its 4 texels. `invResourceSize` must be
// 1 / the size of the sampled texture ( static resolution ).
struct HistoryFilter {
    float4 w;
    float w4, sum;
    float2 tc;
    int ox, oy;
    bool bicubic;
    float4 custom;
    NRD_DEV HistoryFilter(float2 samplePos, float2 /*invResourceSize*/, float4 customWeights, bool useBicubic) {
        const float S = 0.5f;
        float2 centerPos = floor2(samplePos - 0.5f) + 0.5f;
        float2 f = saturate(samplePos - centerPos);
        float2 w0 = f * (f * (-S * f + 2.0f * S) - S);
        float2 w1 = f * (f * ((2.0f - S) * f - (3.0f - S))) + 1.0f;
        float2 w2 = f * (f * (-(2.0f - S) * f + (3.0f - 2.0f * S)) + S);
        float2 w3 = f * (f * (S * f - S));
        float2 w12 = w1 + w2;
        tc = w2 / w12;
        w = useBicubic ? make_float4(w12.x * w0.y, w0.x * w12.y, w12.x * w12.y, w3.x * w12.y) : customWeights;
        w4 = useBicubic ? w12.x * w3.y : 0.0f;
        sum = sum4(w) + w4;
        ox = (int)centerPos.x;
        oy = (int)centerPos.y;
        bicubic = useBicubic;
        custom = customWeights;
    }
    template <class TEX> NRD_DEV auto color(const TEX& tex) const -> decltype(tex.fetchClamped(0, 0)) {
        const int x0 = tex.cx(ox), x1 = tex.cx(ox + 1), y0 = tex.cy(oy), y1 = tex.cy(oy + 1);
        decltype(tex.fetchClamped(0, 0)) c;
#ifdef HF_FETCH_ALL
        {
            // all 12 texels are requested before the filter is chosen ( see the RGBA16F overload )
            using T = decltype(tex.fetchClamped(0, 0));
            const int xm = tex.cx(ox - 1), x2 = tex.cx(ox + 2), ym = tex.cy(oy - 1), y2 = tex.cy(oy + 2);
            const T t00 = tex.fetch(x0, y0), t10 = tex.fetch(x1, y0), t01 = tex.fetch(x0, y1), t11 = tex.fetch(x1, y1);
            const T a0 = tex.fetch(x0, ym), a1 = tex.fetch(x1, ym), b0 = tex.fetch(xm, y0), b1 = tex.fetch(xm, y1);
            const T d0 = tex.fetch(x2, y0), d1 = tex.fetch(x2, y1), e0 = tex.fetch(x0, y2), e1 = tex.fetch(x1, y2);
            if (bicubic) {
                c = lerp(a0, a1, tc.x) * w.x;
                c += lerp(b0, b1, tc.y) * w.y;
                c += lerp(lerp(t00, t10, tc.x), lerp(t01, t11, tc.x), tc.y) * w.z;
                c += lerp(d0, d1, tc.y) * w.w;
                c += lerp(e0, e1, tc.x) * w4;
            } else {
                c = t00 * w.x;
                c += t10 * w.y;
                c += t01 * w.z;
                c += t11 * w.w;
            }
            return sum < 0.0001f ? c * 0.0f : c / sum;
        }
#endif
        if (bicubic) {
            const int xm = tex.cx(ox - 1), x2 = tex.cx(ox + 2), ym = tex.cy(oy - 1), y2 = tex.cy(oy + 2);
            c = lerp(tex.fetch(x0, ym), tex.fetch(x1, ym), tc.x) * w.x;
            c += lerp(tex.fetch(xm, y0), tex.fetch(xm, y1), tc.y) * w.y;
            c += lerp(lerp(tex.fetch(x0, y0), tex.fetch(x1, y0), tc.x), lerp(tex.fetch(x0, y1), tex.fetch(x1, y1), tc.x), tc.y) * w.z;
            c += lerp(tex.fetch(x2, y0), tex.fetch(x2, y1), tc.y) * w.w;
            c += lerp(tex.fetch(x0, y2), tex.fetch(x1, y2), tc.x) * w4;
        } else {
            c = tex.fetch(x0, y0) * w.x;
            c += tex.fetch(x1, y0) * w.y;
            c += tex.fetch(x0, y1) * w.z;
            c += tex.fetch(x1, y1) * w.w;
        }
        return sum < 0.0001f ? c * 0.0f : c / sum;
    }
    // The filter on the signal / fast-history texture of an occlusion mode: same taps through the format-polymorphic accessors ( A = Sig< MODE > or FastSig< MODE > )
    template <class A, class TEX> NRD_DEV auto colorAny(const TEX& tex) const -> decltype(A::fetch(tex, 0, 0)) {
        const int x0 = tex.cx(ox), x1 = tex.cx(ox + 1), y0 = tex.cy(oy), y1 = tex.cy(oy + 1);
        decltype(A::fetch(tex, 0, 0)) c;
        if (bicubic) {
            const int xm = tex.cx(ox - 1), x2 = tex.cx(ox + 2), ym = tex.cy(oy - 1), y2 = tex.cy(oy + 2);
            c = lerp(A::fetch(tex, x0, ym), A::fetch(tex, x1, ym), tc.x) * w.x;
            c += lerp(A::fetch(tex, xm, y0), A::fetch(tex, xm, y1), tc.y) * w.y;
            c += lerp(lerp(A::fetch(tex, x0, y0), A::fetch(tex, x1, y0), tc.x), lerp(A::fetch(tex, x0, y1), A::fetch(tex, x1, y1), tc.x), tc.y) * w.z;
            c += lerp(A::fetch(tex, x2, y0), A::fetch(tex, x2, y1), tc.y) * w.w;
            c += lerp(A::fetch(tex, x0, y2), A::fetch(tex, x1, y2), tc.x) * w4;
        } else {
            c = A::fetch(tex, x0, y0) * w.x;
            c += A::fetch(tex, x1, y0) * w.y;
            c += A::fetch(tex, x0, y1) * w.z;
            c += A::fetch(tex, x1, y1) * w.w;
        }
        return sum < 0.0001f ? c * 0.0f : c / sum;
    }
    template <class A> NRD_DEV float bilinearAny(const TexR16F& tex) const {
        float c = A::load(tex, ox, oy) * custom.x;
        c += A::load(tex, ox + 1, oy) * custom.y;
        c += A::load(tex, ox, oy + 1) * custom.z;
        c += A::load(tex, ox + 1, oy + 1) * custom.w;
        float s = sum4(custom);
        return s < 0.0001f ? 0.0f : c / s;
    }
    // The same filter on an RGBA16F texture with Blackwell's packed fp32x2 pipe: a texel decodes into two register pairs ( .xy, .zw ) for free, so
    // every lerp / weight / sum is one FFMA2-class instruction per pair — half the issue slots of the four-channel scalar form above.
    struct Px { P2 lo, hi; };
    static NRD_DEV Px px(const TexRGBA16F& tex, int x, int y) {
        const uint2 raw = tex.fetchRaw(x, y);
        return {P2(__half22float2(*reinterpret_cast<const __half2*>(&raw.x))), P2(__half22float2(*reinterpret_cast<const __half2*>(&raw.y)))};
    }
    static NRD_DEV Px lerpPx(Px a, Px b, float t) { return {fma2(b.lo - a.lo, t, a.lo), fma2(b.hi - a.hi, t, a.hi)}; }
    NRD_DEV float4 color(const TexRGBA16F& tex) const {
        const int x0 = tex.cx(ox), x1 = tex.cx(ox + 1), y0 = tex.cy(oy), y1 = tex.cy(oy + 1);
        Px c;
#ifdef HF_FETCH_ALL
        {
            // The 12 texels of the Catmull-Rom footprint are requested BEFORE the filter is chosen: with the loads inside the two branches nothing of a later
            // fetch could be issued until the branch of the one before it had been taken ( the kernels that call this wait on memory latency ). The bilinear
            // fallback ( disoccluded footprints, a few percent of the pixels ) reads 8 texels it does not use; the arithmetic of both branches is unchanged.
            const int xm = tex.cx(ox - 1), x2 = tex.cx(ox + 2), ym = tex.cy(oy - 1), y2 = tex.cy(oy + 2);
            const uint2 r00 = tex.fetchRaw(x0, y0), r10 = tex.fetchRaw(x1, y0), r01 = tex.fetchRaw(x0, y1), r11 = tex.fetchRaw(x1, y1);
            const uint2 ra0 = tex.fetchRaw(x0, ym), ra1 = tex.fetchRaw(x1, ym), rb0 = tex.fetchRaw(xm, y0), rb1 = tex.fetchRaw(xm, y1);
            const uint2 rd0 = tex.fetchRaw(x2, y0), rd1 = tex.fetchRaw(x2, y1), re0 = tex.fetchRaw(x0, y2), re1 = tex.fetchRaw(x1, y2);
            auto dec = [](uint2 raw) -> Px { return {P2(__half22float2(*reinterpret_cast<const __half2*>(&raw.x))), P2(__half22float2(*reinterpret_cast<const __half2*>(&raw.y)))}; };
            if (bicubic) {
                Px t = lerpPx(dec(ra0), dec(ra1), tc.x);
                c = {t.lo * w.x, t.hi * w.x};
                t = lerpPx(dec(rb0), dec(rb1), tc.y);
                c = {fma2(t.lo, w.y, c.lo), fma2(t.hi, w.y, c.hi)};
                t = lerpPx(lerpPx(dec(r00), dec(r10), tc.x), lerpPx(dec(r01), dec(r11), tc.x), tc.y);
                c = {fma2(t.lo, w.z, c.lo), fma2(t.hi, w.z, c.hi)};
                t = lerpPx(dec(rd0), dec(rd1), tc.y);
                c = {fma2(t.lo, w.w, c.lo), fma2(t.hi, w.w, c.hi)};
                t = lerpPx(dec(re0), dec(re1), tc.x);
                c = {fma2(t.lo, w4, c.lo), fma2(t.hi, w4, c.hi)};
            } else {
                Px t = dec(r00);
                c = {t.lo * w.x, t.hi * w.x};
                t = dec(r10);
                c = {fma2(t.lo, w.y, c.lo), fma2(t.hi, w.y, c.hi)};
                t = dec(r01);
                c = {fma2(t.lo, w.z, c.lo), fma2(t.hi, w.z, c.hi)};
                t = dec(r11);
                c = {fma2(t.lo, w.w, c.lo), fma2(t.hi, w.w, c.hi)};
            }
            const float k = sum < 0.0001f ? 0.0f : 1.0f / sum;
            c = {c.lo * k, c.hi * k};
            return make_float4(c.lo.a(), c.lo.b(), c.hi.a(), c.hi.b());
        }
#endif
        if (bicubic) {
            const int xm = tex.cx(ox - 1), x2 = tex.cx(ox + 2), ym = tex.cy(oy - 1), y2 = tex.cy(oy + 2);
            Px t = lerpPx(px(tex, x0, ym), px(tex, x1, ym), tc.x);
            c = {t.lo * w.x, t.hi * w.x};
            t = lerpPx(px(tex, xm, y0), px(tex, xm, y1), tc.y);
            c = {fma2(t.lo, w.y, c.lo), fma2(t.hi, w.y, c.hi)};
            t = lerpPx(lerpPx(px(tex, x0, y0), px(tex, x1, y0), tc.x), lerpPx(px(tex, x0, y1), px(tex, x1, y1), tc.x), tc.y);
            c = {fma2(t.lo, w.z, c.lo), fma2(t.hi, w.z, c.hi)};
            t = lerpPx(px(tex, x2, y0), px(tex, x2, y1), tc.y);
            c = {fma2(t.lo, w.w, c.lo), fma2(t.hi, w.w, c.hi)};
            t = lerpPx(px(tex, x0, y2), px(tex, x1, y2), tc.x);
            c = {fma2(t.lo, w4, c.lo), fma2(t.hi, w4, c.hi)};
        } else {
            Px t = px(tex, x0, y0);
            c = {t.lo * w.x, t.hi * w.x};
            t = px(tex, x1, y0);
            c = {fma2(t.lo, w.y, c.lo), fma2(t.hi, w.y, c.hi)};
            t = px(tex, x0, y1);
            c = {fma2(t.lo, w.z, c.lo), fma2(t.hi, w.z, c.hi)};
            t = px(tex, x1, y1);
            c = {fma2(t.lo, w.w, c.lo), fma2(t.hi, w.w, c.hi)};
        }
        const float k = sum < 0.0001f ? 0.0f : 1.0f / sum;
        c = {c.lo * k, c.hi * k};
        return make_float4(c.lo.a(), c.lo.b(), c.hi.a(), c.hi.b());
    }
    // _BilinearFilterWithCustomWeights_Color on a four-channel texture ( the SH history, REBLUR_Common.hlsli:351-371 )
    NRD_DEV float4 bilinear4(const TexRGBA16F& tex) const {
        float4 c = tex.load(ox, oy) * custom.x;
        c += tex.load(ox + 1, oy) * custom.y;
        c += tex.load(ox, oy + 1) * custom.z;
        c += tex.load(ox + 1, oy + 1) * custom.w;
        float s = sum4(custom);
        return s < 0.0001f ? f4(0.0f) : c / s;
    }
    NRD_DEV float bilinear(const TexR16F& tex) const {
        float c = tex.load(ox, oy) * custom.x;
        c += tex.load(ox + 1, oy) * custom.y;
        c += tex.load(ox, oy + 1) * custom.z;
        c += tex.load(ox + 1, oy + 1) * custom.w;
        float s = sum4(custom);
        return s < 0.0001f ? 0.0f : c / s;
    }
};

// ---- launch parameter blocks (member order = shader register order = DispatchDesc::resources order) ----------
struct ClassifyTilesParams {
    TexR32F inViewZ;
    TexTiles outTiles;
};
// REBLUR_Validation.resources.hlsli:22-34; bound by format, not by type: data1 is RG8 or R8, data2 R32_UINT / R8_UINT ( or data1 again for the occlusion denoisers ),
// the lobe inputs whatever the denoiser's inputs are, OUT_VALIDATION "RGBA8+"
struct ReblurValidationParams {
    TexNR normalRoughness; TexR32F viewZ; TexView mv, data1, data2, diff, spec, out;
};
struct SplitScreenParams {
    TexR32F viewZ; TexRGBA16F inDiff, inSpec;
    TexRGBA16F outDiff, outSpec;
    TexRGBA16F inDiffSh, inSpecSh, outDiffSh, outSpecSh;  // NRD_MODE = SH only ( unbound = null views otherwise ): the second RGBA16F of each lobe
};
struct HitDistReconstructionParams {
    TexTiles tiles; TexNR normalRoughness; TexR32F viewZ; TexRGBA16F inDiff, inSpec;
    TexRGBA16F outDiff, outSpec;
};
struct PrePassParams {
    TexTiles tiles; TexNR normalRoughness; TexR32F viewZ; TexRGBA16F inDiff, inSpec;
    TexRGBA16F outDiff, outSpec; TexR16F outSpecHitDistForTracking;
    TexRGBA16F inDiffSh, inSpecSh, outDiffSh, outSpecSh;  // NRD_MODE = SH only
    TexGeom geom;                                          // not a shader binding: the executor's geometry plane of ( normalRoughness, viewZ )
};
// what the geometry plane is decoded from, and the rows it has to cover ( a strip plus the reach of its taps )
struct GeometryPlaneParams {
    TexNR normalRoughness; TexR32F viewZ; TexGeom out;
};
// `data1R8` / `data2R8` / `outData1R8` / `outData2R8`: the single-lobe formats of data1 / data2 (bound instead of the two-lobe view)
struct BlurParams {
    TexTiles tiles; TexNR normalRoughness; TexR32F viewZ; TexRG8 data1; TexRGBA16F inDiff, inSpec;
    TexR32F outViewZ; TexRGBA16F outDiff, outSpec;
    TexR8 data1R8;
    TexRGBA16F inDiffSh, inSpecSh, outDiffSh, outSpecSh;  // NRD_MODE = SH only
    TexGeom geom;
};
struct PostBlurParams {
    TexTiles tiles; TexNR normalRoughness; TexRG8 data1; TexR32F viewZ; TexRGBA16F inDiff, inSpec;
    TexNR outNormalRoughness; TexRGBA16F outDiff, outSpec;
    TexR16U outInternalData; TexRGBA16F outDiffCopy, outSpecCopy;  // only bound when TEMPORAL_STABILIZATION = 0
    TexR8 data1R8;
    TexRGBA16F inDiffSh, inSpecSh, outDiffSh, outSpecSh, outDiffShCopy, outSpecShCopy;  // NRD_MODE = SH only ( out*Sh = the SH history )
    TexGeom geom;
};
struct TemporalAccumulationParams {
    TexTiles tiles; TexNR normalRoughness; TexR32F viewZ; TexRGBA16F mv; TexR32F prevViewZ; TexNR prevNormalRoughness; TexR16U prevInternalData;
    TexAnyX disocclusionThresholdMix, diffConfidence, specConfidence;  // dummies (IN_VIEWZ) unless the optional inputs are enabled
    TexRGBA16F inDiff, inSpec, historyDiff, historySpec; TexR16F historyDiffFast, historySpecFast, prevSpecHitDistForTracking, inSpecHitDistForTracking;
    TexRG8 outData1; TexRGBA16F outDiff, outSpec; TexR16F outDiffFast, outSpecFast, outSpecHitDistForTracking; TexR32U outData2;
    TexR8 outData1R8; TexR8U outData2R8;
    TexRGBA16F inDiffSh, inSpecSh, historyDiffSh, historySpecSh, outDiffSh, outSpecSh;  // NRD_MODE = SH only
};
struct HistoryFixParams {
    TexTiles tiles; TexNR normalRoughness; TexRG8 data1; TexR32F viewZ; TexRGBA16F inDiff, inSpec; TexR16F inDiffFast, inSpecFast, specHitDistForTracking;
    TexRGBA16F outDiff, outSpec; TexR16F outDiffFast, outSpecFast;
    TexR8 data1R8;
    TexRGBA16F inDiffSh, inSpecSh, outDiffSh, outSpecSh;  // NRD_MODE = SH only
};
struct TemporalStabilizationParams {
    TexTiles tiles; TexNR normalRoughness; TexR32F viewZ; TexRG8 data1; TexR32U data2; TexR16F specHitDistForTracking; TexRGBA16F inDiff, inSpec;
    TexR16F historyDiffLuma, historySpecLuma;
    TexRGBA16F mv; TexR16U outInternalData; TexRGBA16F outDiff, outSpec; TexR16F outDiffLuma, outSpecLuma;
    TexR8 data1R8; TexR8U data2R8;
    TexRGBA16F inDiffSh, inSpecSh, outDiffSh, outSpecSh;  // NRD_MODE = SH only ( in*Sh = the SH history )
};

}  // namespace nrdk
