// SIGMA_SHADOW helpers and launch-parameter blocks.
// Reference: External/NRD/Shaders/SIGMA_Common.hlsli:13-95, SIGMA_Config.hlsli:11-42 (switches at their defaults:
// 5x5 radius-estimation and temporal kernels, sparse blur, screen-space sampling, early out in TS, CatRom history).
#pragma once
#include "../../../include/nrd_b200.h"
#include "reblur_common.cuh"  // HistoryFilter (Common.hlsli:604-658) is shared with REBLUR

namespace nrdk {

using nrdb::SigmaConstants;

constexpr float SIGMA_MAX_PIXEL_RADIUS = 32.0f;
constexpr float SIGMA_TS_SIGMA_SCALE = 3.0f;
constexpr float SIGMA_MAX_ACCUM_FRAME_NUM = 7.0f;
constexpr float SIGMA_DISOCCLUSION_THRESHOLD = 0.02f;  // NRD_DISOCCLUSION_THRESHOLD, Common.hlsli:63
constexpr float NRD_FP16_MAX = 65504.0f;
constexpr int SIGMA_BORDER = 2;  // Blur and TemporalStabilization both use the 5x5 neighbourhood

NRD_DEV float sigmaUnpackViewZ(const SigmaConstants& cb, float z) { return fabsf(z * cb.viewZScale); }
NRD_DEV bool sigmaInRange(const SigmaConstants& cb, float z) { return z < cb.denoisingRange; }
NRD_DEV bool sigmaIsLit(float p) { return p >= NRD_FP16_MAX; }
NRD_DEV float sigmaPackShadow(float s) { return sqrt01(s); }
NRD_DEV float4 sigmaPackShadow(float4 s) { return make_float4(sqrt01(s.x), sqrt01(s.y), sqrt01(s.z), sqrt01(s.w)); }

// SIGMA_TYPE (SIGMA_Config.hlsli:37-42): float for SIGMA_SHADOW, float4 = { shadow, translucency.rgb } for SIGMA_SHADOW_TRANSLUCENCY
template <bool TRANSLUCENCY> struct SigmaSignal;
template <> struct SigmaSignal<false> {
    using T = float;
    using Tex = TexR8;
    static constexpr nrd::Format format = nrd::Format::R8_UNORM;
    static NRD_DEV T splat(float v) { return v; }
    static NRD_DEV float x(T s) { return s; }
    static NRD_DEV T stdDev(T m1, T m2) { return nrdk::stdDev(m1, m2); }
    static NRD_DEV T clamp(T v, T lo, T hi) { return fminf(fmaxf(v, lo), hi); }
};
template <> struct SigmaSignal<true> {
    using T = float4;
    using Tex = TexRGBA8;
    static constexpr nrd::Format format = nrd::Format::RGBA8_UNORM;
    static NRD_DEV T splat(float v) { return make_float4(v, v, v, v); }
    static NRD_DEV float x(T s) { return s.x; }
    static NRD_DEV T stdDev(T m1, T m2) { return make_float4(nrdk::stdDev(m1.x, m2.x), nrdk::stdDev(m1.y, m2.y), nrdk::stdDev(m1.z, m2.z), nrdk::stdDev(m1.w, m2.w)); }
    static NRD_DEV T clamp(T v, T lo, T hi) {
        return make_float4(fminf(fmaxf(v.x, lo.x), hi.x), fminf(fmaxf(v.y, lo.y), hi.y), fminf(fmaxf(v.z, lo.z), hi.z), fminf(fmaxf(v.w, lo.w), hi.w));
    }
};
NRD_DEV float sigmaBothLitOrUnlit(float p1, float p2) { return ((p1 == 0.0f) == (p2 == 0.0f)) ? 1.0f : 0.0f; }
// GetKernelRadiusInPixels: fminf / fmaxf are IEEE minNum / maxNum like the HLSL intrinsics (0 / 0 -> lower bound)
NRD_DEV float sigmaKernelRadiusInPixels(float hitDist, float unprojectZ, float scale = 1.0f) {
    float unclamped = hitDist / unprojectZ * scale;
    float minRadius = fminf(unclamped, 2.0f);
    return fminf(fmaxf(unclamped, minRadius), SIGMA_MAX_PIXEL_RADIUS);
}

// TextureCubic( gIn_Tiles, uv ).y ( SIGMA_Common.hlsli:46-95 ): the reference builds the uniform cubic B-spline out of four bilinear taps placed between
// texel pairs; the same sum taken directly over the 4 x 4 texels ( clamped to the edge, like the sampler ) with separable weights is a third of the
// instructions and equal to rounding ( checked against the four-tap form: max difference 2e-15 in double precision ). Only the .y channel is read.
NRD_DEV void sigmaBsplineWeights(float f, float w[4]) {
    const float f2 = f * f, f3 = f2 * f, g = 1.0f - f;
    w[0] = g * g * g;
    w[1] = 3.0f * f3 - 6.0f * f2 + 4.0f;
    w[2] = -3.0f * f3 + 3.0f * f2 + 3.0f * f + 1.0f;
    w[3] = f3;   // ( x 1/6 per axis: applied once at the end )
}
NRD_DEV float sigmaTileValue(const TexRG8& tiles, float2 uv) {
    const float tx = uv.x * (float)tiles.w - 0.5f, ty = uv.y * (float)tiles.h - 0.5f;
    const int ix = __float2int_rd(tx), iy = __float2int_rd(ty);
    float wx[4], wy[4];
    sigmaBsplineWeights(tx - (float)ix, wx);
    sigmaBsplineWeights(ty - (float)iy, wy);
    int xs[4];
#pragma unroll
    for (int i = 0; i < 4; i++) xs[i] = clampi(ix - 1 + i, 0, tiles.w - 1) * 2 + 1;
    float sum = 0.0f;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const uint8_t* row = tiles.data + (size_t)(clampi(iy - 1 + j, 0, tiles.h - 1) * tiles.pitch) * 2u;
        float r = 0.0f;
#pragma unroll
        for (int i = 0; i < 4; i++) r += wx[i] * (float)__ldg(row + xs[i]);
        sum += wy[j] * r;
    }
    return sum * (1.0f / (36.0f * 255.0f));
}

// ---- launch parameter blocks (member order = shader register order = DispatchDesc::resources order) ----------
struct SigmaClassifyTilesParams { TexR32F viewZ; TexR16F penumbra; TexRGBA8 translucency; TexRGBA8 outTiles; };
struct SigmaSmoothTilesParams { TexRGBA8 tiles; TexRG8 outTiles; };
// `ST` = SigmaSignal<TRANSLUCENCY>::Tex: R8 shadow or RGBA8 shadow + translucency
template <class ST> struct SigmaCopyParams { TexRG8 tiles; ST history; TexR32U historyLength; ST outHistory; TexR32U outHistoryLength; };
template <class ST> struct SigmaBlurParams {
    TexR32F viewZ; TexNR normalRoughness; TexR16F penumbra; TexRG8 tiles; ST shadow; TexR16F outPenumbra; ST outShadow;
    // the Copy pass riding along ( first blur pass, contexts only ): previous output + history length -> the transient copies temporal stabilization reads
    ST copyHistory; TexR32U copyHistoryLength; ST copyOutHistory; TexR32U copyOutHistoryLength; int copy;
};
template <class ST> struct SigmaTemporalStabilizationParams {
    TexR32F viewZ; TexRGBA16F mv; TexR16F penumbra; ST shadow; ST history; TexR32U historyLength; TexRG8 tiles;
    ST outShadow; TexR32U outHistoryLength;
};
template <class ST> struct SigmaSplitScreenParams { TexR32F viewZ; TexR16F penumbra; ST translucency; ST outShadow; };

}  // namespace nrdk
