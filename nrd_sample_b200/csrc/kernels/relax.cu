// RELAX_DIFFUSE_SPECULAR_SH and RELAX_DIFFUSE_SPECULAR passes on sm_100a (NRD_SIGNAL = BOTH, NRD_MODE = SH / RADIANCE, template <bool SH>): ClassifyTiles, PrePass,
// TemporalAccumulation, HistoryFix, HistoryClamping, Copy, AntiFirefly, AtrousSmem, Atrous. One kernel per reference
// dispatch, one thread per pixel, CTA = 32x8 pixels (a 32-pixel row per warp: coalesced RGBA16F rows).
//
// Reference (External/NRD/Shaders): RELAX_ClassifyTiles.cs.hlsl:21-51, RELAX_PrePass.cs.hlsl:21-385,
// RELAX_TemporalAccumulation.cs.hlsl:21-942, RELAX_HistoryFix.cs.hlsl:21-163, RELAX_HistoryClamping.cs.hlsl:21-354,
// RELAX_Copy.cs.hlsl:21-34, RELAX_AntiFirefly.cs.hlsl:21-216, RELAX_AtrousSmem.cs.hlsl:21-484, RELAX_Atrous.cs.hlsl:21-260,
// RELAX_SplitScreen.cs.hlsl:21-62, helpers RELAX_Common.hlsli:11-185. Build switches of the reference's default build
// (NRD_USE_PREV_WORLD_SPACE_MATRIX = 0); checkerboard modes, history-confidence and disocclusion-threshold-mix inputs included.
// History clamping ( 36x20 texels per 32x16 pixels, two rows per thread ), the first a-trous pass ( 36x12 ) and the a-trous passes of strides 2 and 4 stage
// their neighbourhoods in shared memory like the reference — colour-space conversions / normal decode / world positions done once per texel, the raw fp16 planes
// of the tiled a-trous passes through TMA — and so does temporal accumulation for its 3x3 of { normal, hitT } ( 34x10 ); anti-firefly goes through L1. The
// arithmetic follows the shaders statement by statement.
#include <cuda.h>

#define HF_FETCH_ALL   // reblur_common.cuh: HistoryFilter requests the whole footprint before it branches ( measured neutral for REBLUR / SIGMA, -8 us here )

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>

#include "../../../include/nrd_b200.h"
#include "../../../include/nrdcu.h"
#include "../pipeline_key.h"
#include "debug_overlay.cuh"  // HistoryFilter

namespace nrdk {

using nrdb::RelaxConstants;

namespace {

constexpr int BLOCK_W = 32, BLOCK_H = 8;
constexpr float RELAX_NORMAL_ULP = 1.5f / 255.0f;
constexpr float RELAX_MAX_ACCUM_FRAME_NUM = 255.0f;
constexpr float RELAX_ANTILAG_ACCELERATION_AMOUNT_SCALE = 10.0f;
constexpr float NRD_FP16_MAX_F = 65504.0f;

// g_Poisson8 ( Common.hlsli:194-205 ): { offset.xy, distance } and GetGaussianWeight( distance ) = exp( -0.66 d^2 ) ( the fp32 values expf gives ) as compile-time
// constants: the unrolled tap loops fold them into immediates instead of loading the offsets from constant memory and evaluating an exponential per tap
__device__ constexpr float kPoisson8[8][3] = {{-0.4706069f, -0.4427112f, +0.6461146f}, {-0.9057375f, +0.3003471f, +0.9542373f}, {-0.3487388f, +0.4037880f, +0.5335386f},
                                              {+0.1023042f, +0.6439373f, +0.6520134f}, {+0.5699277f, +0.3513750f, +0.6695386f}, {+0.2939128f, -0.1131226f, +0.3149309f},
                                              {+0.7836658f, -0.4208784f, +0.8895339f}, {+0.1564120f, -0.8198990f, +0.8346850f}};
__device__ constexpr float kPoisson8Gauss[8] = {0.7591724991798401f, 0.5482766032218933f, 0.8287158608436584f, 0.755345344543457f,
                                                0.7438870072364807f, 0.9366368055343628f, 0.5931912064552307f, 0.6313964128494263f};

// ---- small helpers ------------------------------------------------------------------------------------------------
NRD_DEV float luminance(float3 x) { return dot(x, make_float3(0.2126f, 0.7152f, 0.0722f)); }
NRD_DEV float3 rgbToYCoCg(float3 x) { return make_float3(dot(x, make_float3(0.25f, 0.5f, 0.25f)), dot(x, make_float3(0.5f, 0.0f, -0.5f)), dot(x, make_float3(-0.25f, 0.5f, -0.25f))); }
NRD_DEV float3 yCoCgToRgb(float3 x) { float t = x.x - x.z; return make_float3(t + x.y, x.x + x.z, t - x.y); }
NRD_DEV float3 min3v(float3 a, float3 b) { return make_float3(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)); }
NRD_DEV float4 max4v(float4 a, float b) { return make_float4(fmaxf(a.x, b), fmaxf(a.y, b), fmaxf(a.z, b), fmaxf(a.w, b)); }
NRD_DEV float4 clamp4v(float4 a, float lo, float hi) { return make_float4(fminf(fmaxf(a.x, lo), hi), fminf(fmaxf(a.y, lo), hi), fminf(fmaxf(a.z, lo), hi), fminf(fmaxf(a.w, lo), hi)); }
NRD_DEV float3 clamp3v(float3 a, float3 lo, float3 hi) { return min3v(max3(a, lo), hi); }
NRD_DEV float3 sqrt3v(float3 a) { return make_float3(sqrtf(a.x), sqrtf(a.y), sqrtf(a.z)); }
NRD_DEV float bayer4x4(uint32_t x, uint32_t y, uint32_t frameIndex) {
    const uint32_t px = x & 3u, py = y & 3u;
    const uint32_t b = ((py & 1u) << 2) | ((px & 1u) << 3) | ((py & 2u) >> 1) | (px & 2u);
    return ((float)((b + frameIndex) & 0xFu) + 0.5f) / 16.0f;
}
NRD_DEV float pow5(float x) { return pow01(1.0f - x, 5.0f); }
NRD_DEV float3 safeNormalize(float3 v) { return v * (1.0f / sqrtf(dot(v, v) + 1e-9f)); }
NRD_DEV float4 unpackPrevNormalRoughness(float4 p) { return f4(safeNormalize(xyz(p) * 2.0f - 1.0f), p.w); }
NRD_DEV float4 packPrevNormalRoughness(float4 nr) { return f4(xyz(nr) * 0.5f + 0.5f, nr.w); }
NRD_DEV float customWeightsFloat(float s00, float s10, float s01, float s11, float4 w) {
    float o = s00 * w.x;
    o += s10 * w.y;
    o += s01 * w.z;
    o += s11 * w.w;
    const float sum = sum4(w);
    return sum < 0.0001f ? 0.0f : o * (1.0f / sum);
}
// SH = false is RELAX_DIFFUSE_SPECULAR ( NRD_MODE = RADIANCE ): the SH1 textures are not bound, every access to them compiles away
template <bool SH, class TEX>
NRD_DEV float3 customWeightsSH(const TEX& t, int x, int y, float4 w) {
    if constexpr (!SH) return f3(0.0f);
    float3 o = xyz(t.load(x, y)) * w.x;
    o += xyz(t.load(x + 1, y)) * w.y;
    o += xyz(t.load(x, y + 1)) * w.z;
    o += xyz(t.load(x + 1, y + 1)) * w.w;
    const float sum = sum4(w);
    return sum < 0.0001f ? f3(0.0f) : o * (1.0f / sum);
}
NRD_DEV float planeDistanceWeight(float3 centerWorldPos, float3 centerNormal, float centerViewZ, float3 sampleWorldPos, float threshold) {
    return fabsf(dot(sampleWorldPos - centerWorldPos, centerNormal)) / centerViewZ > threshold ? 0.0f : 1.0f;
}
NRD_DEV float planeDistanceWeightAtrous(float3 centerWorldPos, float3 centerNormal, float3 sampleWorldPos, float threshold) {
    return fabsf(dot(sampleWorldPos - centerWorldPos, centerNormal)) < threshold ? 1.0f : 0.0f;
}
NRD_DEV float relaxSpecLobeTanHalfAngle(float roughness, float percentOfVolume = 0.75f) {
    roughness = saturate(roughness);
    percentOfVolume = saturate(percentOfVolume);
    return roughness * roughness * percentOfVolume / (1.0f - percentOfVolume + NRD_EPS);
}
NRD_DEV float2 normalWeightParamsAtrous(float roughness, float numFramesInHistory, float confidence, float normalEdgeStoppingRelaxation, float lobeAngleFraction, float lobeAngleSlack) {
    float relaxation = saturate(numFramesInHistory / 5.0f);
    relaxation *= lerp(1.0f, confidence, normalEdgeStoppingRelaxation);
    const float f = 0.9f + 0.1f * relaxation;
    float angle = atanf(relaxSpecLobeTanHalfAngle(roughness, lobeAngleFraction));
    angle *= 10.0f - 9.0f * relaxation;
    angle += lobeAngleSlack;
    angle = fminf(3.14159265358979323846f * 0.5f, angle);
    return make_float2(angle, f);
}
NRD_DEV float specularNormalWeightAtrous(float2 params0, float3 n0, float3 n, float3 v0, float3 v) {
    const float cosa = fminf(dot(n0, n), dot(v0, v));
    float a = acosApproxPositive(cosa);
    a = smoothStep(0.0f, params0.x, a);
    return saturate(1.0f - a * params0.y);
}
NRD_DEV float normalWeightParam2(float roughness, float angleFraction) { return 1.0f / fmaxf(atanf(relaxSpecLobeTanHalfAngle(roughness, angleFraction)), RELAX_NORMAL_ULP); }
NRD_DEV float thinLens(float O, float curvature) { return O / (2.0f * curvature * O + 1.0f); }
NRD_DEV float computeWeight(float x, float px, float py) { return nonExponentialWeight(x, px, py); }
NRD_DEV float strandThickness(float strandThicknessV, float pixelSize) { return saturate(0.5f * pixelSize / (strandThicknessV + NRD_EPS)); }

NRD_DEV float relaxViewZ(const RelaxConstants& cb, float z) { return fabsf(z * cb.viewZScale); }
NRD_DEV bool relaxInRange(const RelaxConstants& cb, float z) { return z < cb.denoisingRange; }
NRD_DEV float3 worldPosFrom(const RelaxConstants& cb, float2 clip, float viewZ, const float* fwd, const float* right, const float* up) {
    const float3 F = make_float3(fwd[0], fwd[1], fwd[2]), R = make_float3(right[0], right[1], right[2]), U = make_float3(up[0], up[1], up[2]);
    if (cb.orthoMode == 0.0f) return viewZ * (F + R * clip.x - U * clip.y);
    return viewZ * F + R * clip.x - U * clip.y;
}
NRD_DEV float3 currentWorldPosClip(const RelaxConstants& cb, float2 clip, float viewZ) { return worldPosFrom(cb, clip, viewZ, cb.frustumForward, cb.frustumRight, cb.frustumUp); }
NRD_DEV float3 currentWorldPosPixel(const RelaxConstants& cb, int px, int py, float viewZ) {
    const float2 clip = (make_float2((float)px, (float)py) + 0.5f) * make_float2(cb.rectSizeInv[0], cb.rectSizeInv[1]) * 2.0f - 1.0f;
    return currentWorldPosClip(cb, clip, viewZ);
}
NRD_DEV float3 previousWorldPosClip(const RelaxConstants& cb, float2 clip, float viewZ) { return worldPosFrom(cb, clip, viewZ, cb.prevFrustumForward, cb.prevFrustumRight, cb.prevFrustumUp); }
NRD_DEV float3 previousWorldPosPixel(const RelaxConstants& cb, int px, int py, float viewZ) {
    const float2 clip = (make_float2((float)px, (float)py) + 0.5f) * (f2(1.0f) / make_float2(cb.rectSizePrev[0], cb.rectSizePrev[1])) * 2.0f - 1.0f;
    return previousWorldPosClip(cb, clip, viewZ);
}
NRD_DEV float2 clampUvToViewport(const RelaxConstants& cb, float2 uv) {
    const float2 scale = make_float2(cb.resolutionScale[0], cb.resolutionScale[1]);
    return min2(uv * scale, scale - 0.5f * make_float2(cb.resourceSizeInv[0], cb.resourceSizeInv[1]));
}
NRD_DEV float4 gatherR32(const TexR32F& t, int x0, int y0) { return make_float4(t.fetchClamped(x0, y0), t.fetchClamped(x0 + 1, y0), t.fetchClamped(x0, y0 + 1), t.fetchClamped(x0 + 1, y0 + 1)); }
template <class TEX> NRD_DEV float4 gatherR8(const TEX& t, int x0, int y0) { return make_float4(t.fetchClamped(x0, y0), t.fetchClamped(x0 + 1, y0), t.fetchClamped(x0, y0 + 1), t.fetchClamped(x0 + 1, y0 + 1)); }
template <class TEX> NRD_DEV float4 gatherR16(const TEX& t, int x0, int y0) { return make_float4(t.fetchClamped(x0, y0), t.fetchClamped(x0 + 1, y0), t.fetchClamped(x0, y0 + 1), t.fetchClamped(x0 + 1, y0 + 1)); }
template <bool SH, class TEX> NRD_DEV void storeSh(const TEX& t, int x, int y, float3 v) { if constexpr (SH) t.store(x, y, f4(v, 0.0f)); }
template <bool SH, class TEX> NRD_DEV float3 loadSh(const TEX& t, int x, int y) { if constexpr (SH) return xyz(t.load(x, y)); else return f3(0.0f); }
template <bool SH, class TEX> NRD_DEV float3 sampleNearestSh(const TEX& t, float2 uv) { if constexpr (SH) return xyz(t.sampleNearest(uv)); else return f3(0.0f); }
template <bool SH, class TEX> NRD_DEV float3 fetchSh(const TEX& t, int x, int y) { if constexpr (SH) return xyz(t.fetch(x, y)); else return f3(0.0f); }   // in-bounds coordinates

// ---- NRD_SIGNAL ( DIFF / SPEC / BOTH ) ------------------------------------------------------------------------------------------
// RELAX_DIFFUSE( _SH ) and RELAX_SPECULAR( _SH ) are the two-lobe shaders with the other lobe's `#if( NRD_DIFF )` / `#if( NRD_SPEC )` blocks removed
// ( the only two-lobe expression is TA's history cap, handled where it occurs ). The kernels get that for free from the type system: the
// parameter blocks are templates on SIGNAL, and the texture members of a lobe the denoiser does not have become DeadTex — same layout, every
// read is a compile-time zero, every write a no-op — so the compiler removes the whole lobe ( loads, taps, weights, stores ) from that instantiation.
template <class T> struct DeadTex : TexView {
    using R = decltype(std::declval<const T&>().load(0, 0));
    NRD_DEV R load(int, int) const { return R{}; }
    NRD_DEV R fetch(int, int) const { return R{}; }
    NRD_DEV R fetchClamped(int, int) const { return R{}; }
    NRD_DEV R sampleNearest(float2) const { return R{}; }
    NRD_DEV R sampleLinear(float2) const { return R{}; }
    NRD_DEV void store(int, int, R) const {}
    NRD_DEV int cx(int) const { return 0; }
    NRD_DEV int cy(int) const { return 0; }
};
template <> struct DeadTex<TexAnyX> : TexView {  // TexAnyX carries two more words
    uint32_t kind, bytesPerTexel;
    NRD_DEV float load(int, int) const { return 0.0f; }
    NRD_DEV float fetch(int, int) const { return 0.0f; }
    NRD_DEV float sampleLinear(float2) const { return 0.0f; }
};
template <class T, int SIGNAL> using SpecTex = std::conditional_t<(SIGNAL & SIGNAL_SPEC) != 0, T, DeadTex<T>>;
template <class T, int SIGNAL> using DiffTex = std::conditional_t<(SIGNAL & SIGNAL_DIFF) != 0, T, DeadTex<T>>;
static_assert(sizeof(DeadTex<TexRGBA16F>) == sizeof(TexRGBA16F) && sizeof(DeadTex<TexAnyX>) == sizeof(TexAnyX), "dead views keep the layout of the parameter blocks");

// ---- parameter blocks (member order = DispatchDesc::resources order of the two-lobe denoiser; the executor fills the <SIGNAL_BOTH> block) ----
#define S16 SpecTex<TexRGBA16F, SIGNAL>
#define D16 DiffTex<TexRGBA16F, SIGNAL>
struct RelaxClassifyParams { TexR32F viewZ; TexR8 outTiles; };
template <int SIGNAL> struct RelaxPrePassParamsT { TexR8 tiles; TexNR normalRoughness; TexR32F viewZ; S16 spec; D16 diff; S16 specSh; D16 diffSh; S16 outSpec; D16 outDiff; S16 outSpecSh; D16 outDiffSh; };
template <int SIGNAL> struct RelaxTaParamsT {
    TexR8 tiles; TexRGBA16F mv; TexNR normalRoughness; TexR32F viewZ; TexAnyX mixDummy; TexRGBA8 prevNormalRoughness; TexR32F prevViewZ; TexR8 prevHistoryLength, prevMaterialID;
    S16 spec; D16 diff; S16 historySpecFast; D16 historyDiffFast; S16 historySpec; D16 historyDiff; SpecTex<TexR16F, SIGNAL> prevSpecHitDist; SpecTex<TexAnyX, SIGNAL> specConfDummy;
    DiffTex<TexAnyX, SIGNAL> diffConfDummy;
    S16 specSh; D16 diffSh; S16 historySpecShFast; D16 historyDiffShFast; S16 historySpecSh; D16 historyDiffSh;
    TexR8 outHistoryLength; S16 outSpec; D16 outDiff; S16 outSpecFast; D16 outDiffFast; SpecTex<TexR16F, SIGNAL> outSpecHitDist; SpecTex<TexR8, SIGNAL> outSpecReprojectionConfidence;
    S16 outSpecSh; D16 outDiffSh; S16 outSpecShFast; D16 outDiffShFast;
};
template <int SIGNAL> struct RelaxHistoryFixParamsT { TexR8 tiles, historyLength; TexNR normalRoughness; TexR32F viewZ; S16 spec; D16 diff; S16 specSh; D16 diffSh; S16 outSpec; D16 outDiff; S16 outSpecSh; D16 outDiffSh; };
template <int SIGNAL> struct RelaxHistoryClampingParamsT {
    TexR8 tiles; TexR32F viewZ; TexR8 historyLength; S16 specNoisy; D16 diffNoisy; S16 spec; D16 diff; S16 specFast; D16 diffFast; S16 specSh; D16 diffSh; S16 specShFast; D16 diffShFast;
    TexR8 outHistoryLength; S16 outSpec; D16 outDiff; S16 outSpecFast; D16 outDiffFast; S16 outSpecSh; D16 outDiffSh; S16 outSpecShFast; D16 outDiffShFast;
};
template <int SIGNAL> struct RelaxCopyParamsT { S16 spec; D16 diff; S16 outSpec; D16 outDiff; };
template <int SIGNAL> struct RelaxAntiFireflyParamsT { TexR8 tiles; TexNR normalRoughness; TexR32F viewZ; S16 spec; D16 diff; S16 outSpec; D16 outDiff; };
template <int SIGNAL> struct RelaxAtrousParamsT {
    TexR8 tiles, historyLength; TexNR normalRoughness; TexR32F viewZ; S16 spec; D16 diff; SpecTex<TexR8, SIGNAL> specReprojectionConfidence; SpecTex<TexAnyX, SIGNAL> specConfDummy;
    DiffTex<TexAnyX, SIGNAL> diffConfDummy; S16 specSh; D16 diffSh;
    S16 outSpec; D16 outDiff; TexRGBA8 outNormalRoughness; TexR8 outMaterialID; TexR32F outViewZ; S16 outSpecSh; D16 outDiffSh;
};
using RelaxPrePassParams = RelaxPrePassParamsT<SIGNAL_BOTH>;
using RelaxTaParams = RelaxTaParamsT<SIGNAL_BOTH>;
using RelaxHistoryFixParams = RelaxHistoryFixParamsT<SIGNAL_BOTH>;
using RelaxHistoryClampingParams = RelaxHistoryClampingParamsT<SIGNAL_BOTH>;
using RelaxCopyParams = RelaxCopyParamsT<SIGNAL_BOTH>;
using RelaxAntiFireflyParams = RelaxAntiFireflyParamsT<SIGNAL_BOTH>;
using RelaxAtrousParams = RelaxAtrousParamsT<SIGNAL_BOTH>;
// the <SIGNAL> view of the block the executor filled ( identical layout, see DeadTex )
template <template <int> class P, int SIGNAL> const P<SIGNAL>& lobeView(const P<SIGNAL_BOTH>& p, std::integral_constant<int, SIGNAL>) {
    static_assert(sizeof(P<SIGNAL>) == sizeof(P<SIGNAL_BOTH>), "layout");
    return reinterpret_cast<const P<SIGNAL>&>(p);
}

// ---------------------------------------------------------------------------------------------------------------
// One warp per 16x16 tile, 8 tiles per CTA ( tileIsSkyWarp, common.cuh )
__global__ void __launch_bounds__(256) relaxClassifyTilesKernel(const __grid_constant__ RelaxConstants cb, const __grid_constant__ RelaxClassifyParams p, int ctaY0, int tilesW) {
    pdlEntry();
    const int tx = blockIdx.x * 8 + (threadIdx.x >> 5), ty = blockIdx.y + ctaY0;
    if (tx >= tilesW) return;
    const bool allSky = tileIsSkyWarp(p.viewZ, tx, ty, [&](float z) { return !relaxInRange(cb, fabsf(z)); });
    if ((threadIdx.x & 31) == 0) p.outTiles.store(tx, ty, allSky ? 1.0f : 0.0f);
}

// ---------------------------------------------------------------------------------------------------------------
// Confidence-driven relaxation of the a-trous edge stopping ( RELAX_AtrousSmem.cs.hlsl:201-215, 239-251; RELAX_Atrous.cs.hlsl:67-80, 107-119 ):
// returns { r for the normal weights, r for the luminance weight }
template <class TEX> NRD_DEV float2 confidenceDrivenRelaxation(const RelaxConstants& cb, const TEX& confidence, float2 pixelUv) {
    const float relaxation = saturate(cb.confidenceDrivenRelaxationMultiplier * (1.0f - saturate(confidence.sampleLinear(pixelUv))));
    return make_float2(saturate(relaxation * cb.confidenceDrivenNormalEdgeStoppingRelaxation), saturate(relaxation * cb.confidenceDrivenLuminanceEdgeStoppingRelaxation));
}

// RELAX_Common.hlsli:164-165
NRD_DEV float bilateralWeight(float z, float zc) { return linearStep(0.03f, 0.0f, fabsf(z - zc) * (1.0f / fmaxf(z, zc))); }
// ApplyCheckerboardShift ( Common.hlsli:332-342 ) on a pixel-centre position: move a tap to a pixel that was traced this frame
NRD_DEV float2 applyCheckerboardShift(float2 pos, uint32_t mode, int counter, uint32_t frameIndex) {
    const uint32_t checkerboard = (((uint32_t)(pos.x + 16384.0f) ^ (uint32_t)(pos.y + 16384.0f)) ^ frameIndex) & 1u;
    const float shift = (counter & 1) == 0 ? -1.0f : 1.0f;
    pos.x += (checkerboard != mode && mode != 2u) ? shift : 0.0f;
    return pos;
}

// One Poisson tap of the pre-pass ( RELAX_PrePass.cs.hlsl:139-176 ): the position is snapped to a pixel centre, so every texture of the pass is point-sampled
// at the SAME integer texel — floor( uv * rectSize ), clamped to the rect: scaling the snapped uv by gResolutionScale, clamping it to the viewport and
// multiplying by the texture size lands exactly there. Without checkerboarding that texel is computed once ( one F2I per axis instead of two conversions and
// two multiplies per texture ); with it the inputs are addressed through their own uv as in the reference.
struct PrePassTap {
    float2 uv, uvScaled, uvInput;
    int x, y;
    bool inScreen, byTexel;
    NRD_DEV uint32_t fetchRaw(const TexNR& t) const { return byTexel ? t.fetchRaw(x, y) : t.sampleNearestRaw(uvScaled); }
    NRD_DEV float fetch(const TexR32F& t) const { return byTexel ? t.fetch(x, y) : t.sampleNearest(uvScaled); }
    template <class TEX> NRD_DEV float4 fetchInput(const TEX& t) const { return byTexel ? t.fetch(x, y) : t.sampleNearest(uvInput); }
    template <bool SH, class TEX> NRD_DEV float3 fetchInputSh(const TEX& t) const {
        if constexpr (SH) return xyz(byTexel ? t.fetch(x, y) : t.sampleNearest(uvInput));
        else return f3(0.0f);
    }
};
template <bool CB> NRD_DEV PrePassTap prePassTap(const RelaxConstants& cb, float2 pixelUv, float2 rectSize, float2 rectSizeInv, float4 rotator, int i, float blurRadius, uint32_t lobeCheckerboard,
                                                 bool lobeIsCheckerboarded) {
    PrePassTap t;
    float2 uv = pixelUv * rectSize + rotate2(rotator, make_float2(kPoisson8[i][0], kPoisson8[i][1])) * blurRadius;
    if constexpr (!CB) {
        const int kx = __float2int_rd(uv.x), ky = __float2int_rd(uv.y);
        t.uv = make_float2(((float)kx + 0.5f) * rectSizeInv.x, ((float)ky + 0.5f) * rectSizeInv.y);
        t.inScreen = (unsigned)kx < (unsigned)cb.rectSize[0] && (unsigned)ky < (unsigned)cb.rectSize[1];   // IsInScreenNearest: 0 < ( k + 0.5 ) / size < 1
        t.x = clampi(kx, 0, cb.rectSize[0] - 1);
        t.y = clampi(ky, 0, cb.rectSize[1] - 1);
        t.byTexel = true;
        t.uvScaled = t.uvInput = t.uv;   // unused
    } else {
        uv = floor2(uv) + 0.5f;
        uv = applyCheckerboardShift(uv, lobeCheckerboard, i, cb.frameIndex);
        uv = uv * rectSizeInv;
        t.uv = uv;
        t.uvScaled = clampUvToViewport(cb, uv);
        t.uvInput = make_float2(lobeIsCheckerboarded ? t.uvScaled.x * 0.5f : t.uvScaled.x, t.uvScaled.y);
        t.inScreen = isInScreenNearest(uv);
        t.x = t.y = 0;
        t.byTexel = false;
    }
    return t;
}

// CB: checkerboarded inputs ( CheckerboardMode::BLACK / WHITE ): the traced pixels sit in the left half of the input textures
template <bool SH, bool CB, int SIGNAL>
__global__ void __launch_bounds__(BLOCK_W* BLOCK_H) relaxPrePassKernel(const __grid_constant__ RelaxConstants cb, const __grid_constant__ RelaxPrePassParamsT<SIGNAL> p, int ctaY0) {
    pdlEntry();
    const int px = blockIdx.x * BLOCK_W + threadIdx.x, py = (blockIdx.y + ctaY0) * BLOCK_H + threadIdx.y;
    if (p.tiles.load(px >> 4, py >> 4) != 0.0f || px >= cb.rectSize[0] || py >= cb.rectSize[1]) return;
    const float centerViewZ = relaxViewZ(cb, p.viewZ.load(px, py));
    if (!relaxInRange(cb, centerViewZ)) return;

    float centerMaterialID;
    const float4 centerNormalRoughness = unpackNormalRoughness(p.normalRoughness.loadRaw(px, py), centerMaterialID);
    const float3 centerNormal = xyz(centerNormalRoughness);
    const float centerRoughness = centerNormalRoughness.w;
    const float3 centerWorldPos = currentWorldPosPixel(cb, px, py, centerViewZ);
    const float4 rotator = make_float4(cb.rotatorPre[0], cb.rotatorPre[1], cb.rotatorPre[2], cb.rotatorPre[3]);
    const float2 rectSize = make_float2((float)cb.rectSize[0], (float)cb.rectSize[1]), rectSizeInv = make_float2(cb.rectSizeInv[0], cb.rectSizeInv[1]);
    const float2 pixelUv = make_float2((float)px + 0.5f, (float)py + 0.5f) * rectSizeInv;
    const float minRectDim = (float)min(cb.rectSize[0], cb.rectSize[1]);
    const float planeZ = cb.orthoMode == 0.0f ? centerViewZ : 1.0f;

    // Checkerboard resolve weights ( RELAX_PrePass.cs.hlsl:39-72 )
    uint32_t checkerboard = 0;
    int cbX0 = 0, cbX1 = 0;
    float materialID0 = 0.0f, materialID1 = 0.0f;
    float2 resolveWeights = f2(1.0f);
    if (CB) {
        checkerboard = ((uint32_t)(px ^ py) ^ cb.frameIndex) & 1u;
        cbX0 = max(px - 1, 0);
        cbX1 = min(px + 1, cb.rectSize[0] - 1);
        const float viewZ0 = relaxViewZ(cb, p.viewZ.load(cbX0, py)), viewZ1 = relaxViewZ(cb, p.viewZ.load(cbX1, py));
        materialID0 = materialFromRaw(p.normalRoughness.loadRaw(cbX0, py));
        materialID1 = materialFromRaw(p.normalRoughness.loadRaw(cbX1, py));
        resolveWeights = make_float2(bilateralWeight(viewZ0, centerViewZ), bilateralWeight(viewZ1, centerViewZ));
        if (!relaxInRange(cb, viewZ0) || px < 1) resolveWeights.x = 0.0f;
        if (!relaxInRange(cb, viewZ1) || px > cb.rectSize[0] - 2) resolveWeights.y = 0.0f;
        cbX0 >>= 1;
        cbX1 >>= 1;
    }
    // the traced row neighbours of a pixel the checkerboard skipped this frame ( :101-125, 238-262 )
    auto resolve = [&](const auto& tex, const auto& texSh, float minMaterial, float4& illumination, float3& sh) {
        float2 wc = resolveWeights;
        wc.x *= compareMaterials(centerMaterialID, materialID0, minMaterial) ? 1.0f : 0.0f;
        wc.y *= compareMaterials(centerMaterialID, materialID1, minMaterial) ? 1.0f : 0.0f;
        wc = wc * positiveRcp(wc.x + wc.y);
        float4 a = tex.load(cbX0, py), b = tex.load(cbX1, py);
        float3 aSh = loadSh<SH>(texSh, cbX0, py), bSh = loadSh<SH>(texSh, cbX1, py);
        if (wc.x == 0.0f) { a = f4(0.0f); aSh = f3(0.0f); }
        if (wc.y == 0.0f) { b = f4(0.0f); bSh = f3(0.0f); }
        illumination = a * wc.x + b * wc.y;
        sh = aSh * wc.x + bSh * wc.y;
    };

    // ---- diffuse ----
    const bool diffCb = CB && cb.diffCheckerboard != 2u;
    float4 diffuseIllumination = p.diff.load(diffCb ? px >> 1 : px, py);
    float3 diffuseSH = loadSh<SH>(p.diffSh, diffCb ? px >> 1 : px, py);
    if (diffCb && checkerboard != cb.diffCheckerboard) resolve(p.diff, p.diffSh, cb.diffMinMaterial, diffuseIllumination, diffuseSH);
    if (cb.diffBlurRadius > 0.0f) {
        const float frustumSize = pixelRadiusToWorld(cb.unproject, cb.orthoMode, minRectDim, centerViewZ);
        const float hitDist = diffuseIllumination.w == 0.0f ? 1.0f : diffuseIllumination.w;
        float blurRadius = cb.diffBlurRadius * hitDistFactor(hitDist, frustumSize);
        if (diffuseIllumination.w == 0.0f) blurRadius = fmaxf(blurRadius, 1.0f);
        const float normalWeightParam = normalWeightParam2(1.0f, 0.25f * cb.lobeAngleFraction);
        const float2 hitDistanceWeightP = hitDistanceWeightParams(diffuseIllumination.w, 1.0f / 9.0f);
        float weightSum = 1.0f;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const PrePassTap tap = prePassTap<CB>(cb, pixelUv, rectSize, rectSizeInv, rotator, i, blurRadius, cb.diffCheckerboard, diffCb);
            const float2 uv = tap.uv;
            float sampleMaterialID;
            const float3 sampleNormal = xyz(unpackNormalRoughness(tap.fetchRaw(p.normalRoughness), sampleMaterialID));
            const float sampleViewZ = relaxViewZ(cb, tap.fetch(p.viewZ));
            const float3 sampleWorldPos = currentWorldPosClip(cb, uv * 2.0f - 1.0f, sampleViewZ);

            float sampleWeight = tap.inScreen ? 1.0f : 0.0f;
            sampleWeight *= relaxInRange(cb, sampleViewZ) ? 1.0f : 0.0f;
            sampleWeight *= compareMaterials(centerMaterialID, sampleMaterialID, cb.diffMinMaterial) ? 1.0f : 0.0f;
            sampleWeight *= planeDistanceWeight(centerWorldPos, centerNormal, planeZ, sampleWorldPos, cb.depthThreshold);
            sampleWeight *= computeWeight(acosApproxPositive(dot(centerNormal, sampleNormal)), normalWeightParam, 0.0f);

            float4 sampleDiffuse = tap.fetchInput(p.diff);
            if (sampleWeight == 0.0f) sampleDiffuse = f4(0.0f);
            sampleWeight *= lerp(cb.minHitDistanceWeight, 1.0f, exponentialWeight(sampleDiffuse.w, hitDistanceWeightP.x, hitDistanceWeightP.y));
            sampleWeight *= kPoisson8Gauss[i];

            weightSum += sampleWeight;
            diffuseIllumination += sampleDiffuse * sampleWeight;
            float3 sampleSH = tap.template fetchInputSh<SH>(p.diffSh);
            if (sampleWeight == 0.0f) sampleSH = f3(0.0f);
            diffuseSH += sampleSH * sampleWeight;
        }
        diffuseIllumination = diffuseIllumination / weightSum;
        diffuseSH = diffuseSH / weightSum;
    }
    p.outDiff.store(px, py, clamp4v(diffuseIllumination, 0.0f, NRD_FP16_MAX_F));
    storeSh<SH>(p.outDiffSh, px, py, clamp3v(diffuseSH, f3(-NRD_FP16_MAX_F), f3(NRD_FP16_MAX_F)));

    // ---- specular ----
    Rng rng;
    rng.init((uint32_t)px, (uint32_t)py, cb.frameIndex);
    const bool specCb = CB && cb.specCheckerboard != 2u;
    float4 specularIllumination = p.spec.load(specCb ? px >> 1 : px, py);
    float3 specularSH = loadSh<SH>(p.specSh, specCb ? px >> 1 : px, py);
    if (specCb && checkerboard != cb.specCheckerboard) resolve(p.spec, p.specSh, cb.specMinMaterial, specularIllumination, specularSH);
    specularIllumination.w = fmaxf(0.0f, fminf(cb.denoisingRange, specularIllumination.w));
    if (cb.specBlurRadius > 0.0f) {
        const float3 viewVector = cb.orthoMode == 0.0f ? normalize(-centerWorldPos) : make_float3(cb.frustumForward[0], cb.frustumForward[1], cb.frustumForward[2]);
        const float4 D = specularDominantDirectionG2(centerNormal, viewVector, centerRoughness);
        const float NoD = fabsf(dot(centerNormal, xyz(D)));
        const float frustumSize = pixelRadiusToWorld(cb.unproject, cb.orthoMode, minRectDim, centerViewZ);
        const float hitDist = specularIllumination.w == 0.0f ? 1.0f : specularIllumination.w;
        const float smc = specMagicCurve(centerRoughness);
        float blurRadius = cb.specBlurRadius * hitDistFactor(hitDist * NoD, frustumSize) * smc;
        const float lobeRadius = hitDist * NoD * specularLobeTanHalfAngle(centerRoughness, 0.75f);
        const float minBlurRadius = lobeRadius / pixelRadiusToWorld(cb.unproject, cb.orthoMode, 1.0f, centerViewZ + hitDist * D.w);
        blurRadius = fminf(blurRadius, minBlurRadius);
        if (specularIllumination.w == 0.0f) blurRadius = fmaxf(blurRadius, 1.0f);

        const float normalWeightParam = normalWeightParam2(centerRoughness, 0.5f * cb.lobeAngleFraction);
        const float2 hitDistanceWeightP = hitDistanceWeightParams(specularIllumination.w, 1.0f / 9.0f);
        const float2 roughnessWeightP = roughnessWeightParams(centerRoughness, cb.roughnessFraction);
        const float specMinHitDistanceWeight = specularIllumination.w == 0.0f ? 1.0f : cb.minHitDistanceWeight * smc;
        const float specularHitT = specularIllumination.w == 0.0f ? cb.denoisingRange : specularIllumination.w;
        const float NoV = fabsf(dot(centerNormal, viewVector));
        float minHitT = specularHitT == 0.0f ? NRD_INF : specularHitT;
        float weightSum = 1.0f;
        const float roughnessLerp = linearStep(0.5f, 1.0f, centerRoughness);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const PrePassTap tap = prePassTap<CB>(cb, pixelUv, rectSize, rectSizeInv, rotator, i, blurRadius, cb.specCheckerboard, specCb);
            const float2 uv = tap.uv;
            float sampleMaterialID;
            const float4 sampleNormalRoughness = unpackNormalRoughness(tap.fetchRaw(p.normalRoughness), sampleMaterialID);
            const float3 sampleNormal = xyz(sampleNormalRoughness);
            const float sampleViewZ = relaxViewZ(cb, tap.fetch(p.viewZ));

            float sampleWeight = tap.inScreen ? 1.0f : 0.0f;
            sampleWeight *= relaxInRange(cb, sampleViewZ) ? 1.0f : 0.0f;
            sampleWeight *= compareMaterials(centerMaterialID, sampleMaterialID, cb.specMinMaterial) ? 1.0f : 0.0f;
            sampleWeight *= computeWeight(sampleNormalRoughness.w, roughnessWeightP.x, roughnessWeightP.y);
            sampleWeight *= computeWeight(acosApproxPositive(dot(centerNormal, sampleNormal)), normalWeightParam, 0.0f);
            const float3 sampleWorldPos = currentWorldPosClip(cb, uv * 2.0f - 1.0f, sampleViewZ);
            sampleWeight *= planeDistanceWeight(centerWorldPos, centerNormal, planeZ, sampleWorldPos, cb.depthThreshold);

            float4 sampleSpecular = tap.fetchInput(p.spec);
            if (sampleWeight == 0.0f) sampleSpecular = f4(0.0f);
            if (rng.next() < sampleWeight * NoV) minHitT = fminf(minHitT, sampleSpecular.w == 0.0f ? NRD_INF : sampleSpecular.w);

            sampleWeight *= lerp(specMinHitDistanceWeight, 1.0f, exponentialWeight(sampleSpecular.w, hitDistanceWeightP.x, hitDistanceWeightP.y));
            sampleWeight *= kPoisson8Gauss[i];
            const float d = length(sampleWorldPos - centerWorldPos);
            const float t = sampleSpecular.w / (specularIllumination.w + d);
            sampleWeight *= lerp(saturate(t), 1.0f, roughnessLerp);

            weightSum += sampleWeight;
            specularIllumination.x += sampleSpecular.x * sampleWeight;
            specularIllumination.y += sampleSpecular.y * sampleWeight;
            specularIllumination.z += sampleSpecular.z * sampleWeight;
            float3 sampleSH = tap.template fetchInputSh<SH>(p.specSh);
            if (sampleWeight == 0.0f) sampleSH = f3(0.0f);
            specularSH += sampleSH * sampleWeight;
        }
        specularIllumination = f4(xyz(specularIllumination) / weightSum, minHitT == NRD_INF ? 0.0f : minHitT);
        specularSH = specularSH / weightSum;
    }
    p.outSpec.store(px, py, clamp4v(specularIllumination, 0.0f, NRD_FP16_MAX_F));
    storeSh<SH>(p.outSpecSh, px, py, clamp3v(specularSH, f3(-NRD_FP16_MAX_F), f3(NRD_FP16_MAX_F)));
}

// ---------------------------------------------------------------------------------------------------------------
#ifndef RELAX_TA_MIN_BLOCKS
#define RELAX_TA_MIN_BLOCKS 4  // 64 regs: 612 us vs 652 us at 80 regs (3 CTAs) and 714 us at 128 regs (2 CTAs) per 1440p frame
#endif
// OPT: checkerboard resolve speed-up and the application's guide textures ( confidence, threshold mix ); compiled out of the plain kernel
template <bool SH, bool OPT, int SIGNAL>
__global__ void __launch_bounds__(BLOCK_W* BLOCK_H, RELAX_TA_MIN_BLOCKS) relaxTemporalAccumulationKernel(const __grid_constant__ RelaxConstants cb, const __grid_constant__ RelaxTaParamsT<SIGNAL> p, int ctaY0) {
    pdlEntry();
    // Preload( ) of the shader ( RELAX_TemporalAccumulation.cs.hlsl: s_Normal_SpecHitT ): { normal, spec hitT } at the rect-clamped positions of the CTA's 34x10
    // neighbourhood, unpacked ONCE per texel into shared memory like the reference's group-shared tile. Every pixel reads 11 of them ( the 3x3 plus the two
    // curvature neighbours again ); decoding per read was ~12 % of the kernel's instructions.
    constexpr int TA_TILE_W = BLOCK_W + 2, TA_TILE_H = BLOCK_H + 2;
    __shared__ float4 sNormalHitT[TA_TILE_H][TA_TILE_W];
    const int px = blockIdx.x * BLOCK_W + threadIdx.x, py = (blockIdx.y + ctaY0) * BLOCK_H + threadIdx.y;
    const int maxX = cb.rectSize[0] - 1, maxY = cb.rectSize[1] - 1;
    // the CTA covers two 16x16 tiles of one tile row: nothing to do ( and nothing to stage ) when both are sky
    const float skyL = p.tiles.load((px - (int)threadIdx.x) >> 4, py >> 4), skyR = p.tiles.load((px - (int)threadIdx.x + 16) >> 4, py >> 4);
    if (skyL != 0.0f && skyR != 0.0f) return;
    {
        const int baseX = px - (int)threadIdx.x - 1, baseY = py - (int)threadIdx.y - 1;
        for (int i = threadIdx.y * BLOCK_W + threadIdx.x; i < TA_TILE_W * TA_TILE_H; i += BLOCK_W * BLOCK_H) {
            const int sx = i % TA_TILE_W, sy = i / TA_TILE_W;
            const int gx = clampi(baseX + sx, 0, maxX), gy = clampi(baseY + sy, 0, maxY);
            sNormalHitT[sy][sx] = f4(xyz(unpackNormalRoughness(p.normalRoughness.loadRaw(gx, gy))), p.spec.load(gx, gy).w);
        }
    }
    __syncthreads();
    if ((threadIdx.x < 16 ? skyL : skyR) != 0.0f || px >= cb.rectSize[0] || py >= cb.rectSize[1]) return;
    const float currentLinearZ = relaxViewZ(cb, p.viewZ.load(px, py));
    const uint32_t checkerboard = ((uint32_t)(px ^ py) ^ cb.frameIndex) & 1u;
    if (!relaxInRange(cb, currentLinearZ)) return;

    const float2 rectSize = make_float2((float)cb.rectSize[0], (float)cb.rectSize[1]), rectSizeInv = make_float2(cb.rectSizeInv[0], cb.rectSizeInv[1]);
    const float2 rectSizePrev = make_float2(cb.rectSizePrev[0], cb.rectSizePrev[1]), resourceSizeInvPrev = make_float2(cb.resourceSizeInvPrev[0], cb.resourceSizeInvPrev[1]);
    const float2 resolutionScalePrev = rectSizePrev * resourceSizeInvPrev;
    const float3 cameraDelta = make_float3(cb.cameraDelta[0], cb.cameraDelta[1], cb.cameraDelta[2]);
    const float minRectDim = (float)min(cb.rectSize[0], cb.rectSize[1]);
    // ( px + dx, py + dy ) with dx, dy in -1..1: the staged texel ( clamped to the rect while staging, like the per-read clamp it replaces )
    auto preload = [&](int x, int y) -> float4 { return sNormalHitT[y - py + (int)threadIdx.y + 1][x - px + (int)threadIdx.x + 1]; };

    float currentMaterialID;
    const float4 currentNormalRoughness = unpackNormalRoughness(p.normalRoughness.loadRaw(px, py), currentMaterialID);
    const float3 currentNormal = xyz(currentNormalRoughness);
    const float currentRoughness = currentNormalRoughness.w;

    const float3 currentWorldPos = currentWorldPosPixel(cb, px, py, currentLinearZ);
    const float3 fwd = make_float3(cb.frustumForward[0], cb.frustumForward[1], cb.frustumForward[2]);
    const float3 currentViewVector = cb.orthoMode == 0.0f ? currentWorldPos : currentLinearZ * normalize(fwd);
    const float3 V = -normalize(currentViewVector);
    const float NoV = fabsf(dot(currentNormal, V));

    const float2 pixelUv = make_float2((float)px + 0.5f, (float)py + 0.5f) * rectSizeInv;
    const float4 mvRaw = p.mv.load(px, py);
    float3 mv = make_float3(mvRaw.x * cb.mvScale[0], mvRaw.y * cb.mvScale[1], mvRaw.z * cb.mvScale[2]);
    float3 prevWorldPos = currentWorldPos;
    float2 prevUVSMB = pixelUv + make_float2(mv.x, mv.y);
    if (cb.mvScale[3] == 0.0f) {
        if (cb.mvScale[2] == 0.0f) mv.z = affine(cb.worldToViewPrev, currentWorldPos).z - currentLinearZ;
        prevWorldPos = previousWorldPosClip(cb, prevUVSMB * 2.0f - 1.0f, currentLinearZ + mv.z) + cameraDelta;
    } else {
        prevWorldPos = prevWorldPos + mv;
        prevUVSMB = screenUv(cb.worldToClipPrev, prevWorldPos);
    }

    const float3 diffuseIllumination = xyz(p.diff.load(px, py));
    const float3 diffuseSH = loadSh<SH>(p.diffSh, px, py);
    const float4 specularIllumination = p.spec.load(px, py);
    const float3 specularSH = loadSh<SH>(p.specSh, px, py);

    const float hitTM1 = preload(px, py).w;
    float minHitDist3x3 = hitTM1 == 0.0f ? NRD_INF : hitTM1;
    float3 currentNormalAveraged = currentNormal;
#pragma unroll
    for (int i = -1; i <= 1; i++)
#pragma unroll
        for (int j = -1; j <= 1; j++) {
            if (i == 0 && j == 0) continue;
            const float4 n = preload(px + i, py + j);
            minHitDist3x3 = fminf(minHitDist3x3, n.w == 0.0f ? NRD_INF : n.w);
            currentNormalAveraged += xyz(n);
        }
    currentNormalAveraged = currentNormalAveraged / 9.0f;
    const float currentRoughnessModified = modifiedRoughnessFromNormalVariance(currentRoughness, currentNormalAveraged);

    const float specular1stMoment = luminance(xyz(specularIllumination)), specular2ndMoment = specular1stMoment * specular1stMoment;
    const float diffuse1stMoment = luminance(diffuseIllumination), diffuse2ndMoment = diffuse1stMoment * diffuse1stMoment;

    const float smbParallaxInPixels1 = parallaxInPixels(prevWorldPos + cameraDelta, cb.orthoMode == 0.0f ? prevUVSMB : pixelUv, cb.worldToClipPrev, rectSize);
    const float smbParallaxInPixels2 = parallaxInPixels(prevWorldPos - cameraDelta, cb.orthoMode == 0.0f ? pixelUv : prevUVSMB, cb.worldToClip, rectSize);
    const float smbParallaxInPixelsMax = fmaxf(smbParallaxInPixels1, smbParallaxInPixels2), smbParallaxInPixelsMin = fminf(smbParallaxInPixels1, smbParallaxInPixels2);
    const float pixelSize = pixelRadiusToWorld(cb.unproject, cb.orthoMode, 1.0f, currentLinearZ);

    float disocclusionThresholdMix = 0.0f;
    if (currentMaterialID == cb.strandMaterialID) disocclusionThresholdMix = strandThickness(cb.strandThickness, pixelSize);
    if (OPT && cb.hasDisocclusionThresholdMix) disocclusionThresholdMix = p.mixDummy.load(px, py);
    float disocclusionThreshold = lerp(cb.disocclusionThreshold, cb.disocclusionThresholdAlternate, disocclusionThresholdMix);
    if (currentMaterialID == cb.strandMaterialID) disocclusionThreshold = lerp(0.25f, disocclusionThreshold, smoothStep01(smbParallaxInPixelsMax));

    // ---- surface-motion based history (TA:48-236) ----
    float footprintQuality, historyLength, SMBReprojectionFound, prevReflectionHitTSMB;
    float4 prevDiffuseSMB, prevSpecularSMB;
    float3 prevDiffuseSMBResponsive, prevSpecularSMBResponsive, prevDiffuseSH, prevDiffuseResponsiveSH, prevSpecularSMBSH, prevSpecularSMBResponsiveSH;
    {
        const float3 smbNormal = normalize(currentNormalAveraged);
        const float2 prevPixelPosFloat = prevUVSMB * rectSizePrev;
        const float2 fl = floor2(prevPixelPosFloat - 0.5f);
        const int ox = (int)fl.x, oy = (int)fl.y;
        Bilinear bilinear;
        bilinear.origin = fl;
        bilinear.weights = frac2(prevPixelPosFloat - 0.5f);

        auto unpack4 = [&](float4 z) { return make_float4(relaxViewZ(cb, z.x), relaxViewZ(cb, z.y), relaxViewZ(cb, z.z), relaxViewZ(cb, z.w)); };
        const float4 z00 = unpack4(gatherR32(p.prevViewZ, ox - 1, oy - 1)), z10 = unpack4(gatherR32(p.prevViewZ, ox + 1, oy - 1));
        const float4 z01 = unpack4(gatherR32(p.prevViewZ, ox - 1, oy + 1)), z11 = unpack4(gatherR32(p.prevViewZ, ox + 1, oy + 1));
        const float4 m00 = gatherR8(p.prevMaterialID, ox - 1, oy - 1) * 255.0f, m10 = gatherR8(p.prevMaterialID, ox + 1, oy - 1) * 255.0f;
        const float4 m01 = gatherR8(p.prevMaterialID, ox - 1, oy + 1) * 255.0f, m11 = gatherR8(p.prevMaterialID, ox + 1, oy + 1) * 255.0f;

        const float frustumSize = pixelSize * minRectDim;
        const float slopeScale = 1.0f / lerp(lerp(0.05f, 1.0f, NoV), 1.0f, saturate(smbParallaxInPixelsMax / 30.0f));
        float4 thr = f4(saturate(disocclusionThreshold * slopeScale) * frustumSize);
        thr = thr * isInScreenBilinear(fl, rectSizePrev);
        thr = thr - NRD_EPS;

        const float pz = affine(cb.worldToViewPrev, prevWorldPos).z;
        const float minMaterial = fminf(cb.specMinMaterial, cb.diffMinMaterial);
        auto valid = [&](float z, float t, float mat) -> float { return (t >= fabsf(z - pz) ? 1.0f : 0.0f) * (compareMaterials(currentMaterialID, mat, minMaterial) ? 1.0f : 0.0f); };
        const float3 tv0 = make_float3(valid(z00.y, thr.x, m00.y), valid(z00.z, thr.x, m00.z), valid(z00.w, thr.x, m00.w));
        const float3 tv1 = make_float3(valid(z10.x, thr.y, m10.x), valid(z10.z, thr.y, m10.z), valid(z10.w, thr.y, m10.w));
        const float3 tv2 = make_float3(valid(z01.x, thr.z, m01.x), valid(z01.y, thr.z, m01.y), valid(z01.w, thr.z, m01.w));
        const float3 tv3 = make_float3(valid(z11.x, thr.w, m11.x), valid(z11.y, thr.w, m11.y), valid(z11.z, thr.w, m11.z));
        const float3 tvSum = tv0 + tv1 + tv2 + tv3;
        float bicubicFootprintValid = (tvSum.x + tvSum.y + tvSum.z) > 11.5f ? 1.0f : 0.0f;
        float4 bilinearTapsValid = make_float4(tv0.z, tv1.y, tv2.y, tv3.x);

        const float2 uv = (fl + 1.0f) * resourceSizeInvPrev;
        const float3 prevNormalFlat = xyz(unpackPrevNormalRoughness(p.prevNormalRoughness.sampleLinear(uv)));
        if (dot(smbNormal, prevNormalFlat) < 0.0f) {
            bilinearTapsValid = f4(0.0f);
            bicubicFootprintValid = 0.0f;
        }
        const float4 bilinearCustomW = bilinearCustomWeights(bilinear, bilinearTapsValid);
        const bool useBicubic = bicubicFootprintValid > 0.0f;

        const HistoryFilter hf(prevPixelPosFloat, resourceSizeInvPrev, bilinearCustomW, useBicubic);
        prevDiffuseSMB = max4v(hf.color(p.historyDiff), 0.0f);
        prevSpecularSMB = max4v(hf.color(p.historySpec), 0.0f);
        prevDiffuseSMBResponsive = max3(xyz(hf.color(p.historyDiffFast)), f3(0.0f));
        prevSpecularSMBResponsive = max3(xyz(hf.color(p.historySpecFast)), f3(0.0f));

        prevDiffuseSH = customWeightsSH<SH>(p.historyDiffSh, ox, oy, bilinearCustomW);
        prevDiffuseResponsiveSH = customWeightsSH<SH>(p.historyDiffShFast, ox, oy, bilinearCustomW);
        prevSpecularSMBSH = customWeightsSH<SH>(p.historySpecSh, ox, oy, bilinearCustomW);
        prevSpecularSMBResponsiveSH = customWeightsSH<SH>(p.historySpecShFast, ox, oy, bilinearCustomW);

        const float4 prevHistoryLengths = gatherR8(p.prevHistoryLength, ox, oy);
        historyLength = 255.0f * customWeightsFloat(prevHistoryLengths.x, prevHistoryLengths.y, prevHistoryLengths.z, prevHistoryLengths.w, bilinearCustomW);
        const float4 prevHitTs = gatherR16(p.prevSpecHitDist, ox, oy);
        prevReflectionHitTSMB = fmaxf(0.001f, customWeightsFloat(prevHitTs.x, prevHitTs.y, prevHitTs.z, prevHitTs.w, bilinearCustomW));

        SMBReprojectionFound = bicubicFootprintValid > 0.0f ? 2.0f : 1.0f;
        footprintQuality = bicubicFootprintValid > 0.0f ? 1.0f : sum4(bilinearCustomW);
        if (!(bilinearTapsValid.x != 0.0f || bilinearTapsValid.y != 0.0f || bilinearTapsValid.z != 0.0f || bilinearTapsValid.w != 0.0f)) {
            SMBReprojectionFound = 0.0f;
            footprintQuality = 0.0f;
        }
    }

    historyLength = fminf(RELAX_MAX_ACCUM_FRAME_NUM, historyLength + 1.0f);
    const float3 prevFwd = make_float3(cb.prevFrustumForward[0], cb.prevFrustumForward[1], cb.prevFrustumForward[2]);
    const float3 Vprev = cb.orthoMode == 0.0f ? -normalize(prevWorldPos - cameraDelta) : -normalize(prevFwd);
    const float NoVprev = fabsf(dot(currentNormal, Vprev));
    float sizeQuality = (NoVprev + 1e-3f) / (NoV + 1e-3f);
    sizeQuality *= sizeQuality;
    sizeQuality *= sizeQuality;
    footprintQuality *= lerp(0.1f, 1.0f, saturate(sizeQuality + fabsf(cb.orthoMode)));
    if (footprintQuality < 1.0f) historyLength = fmaxf(historyLength * sqrtf(footprintQuality), 1.0f);
    historyLength = cb.resetHistory != 0 ? 1.0f : historyLength;
    // TA:582-588: the cap is the longer of the two lobes' histories, or the only lobe's
    historyLength = fminf(historyLength, 1.0f + (SIGNAL == SIGNAL_BOTH ? fmaxf(cb.diffMaxAccumulatedFrameNum, cb.specMaxAccumulatedFrameNum)
                                                                     : (SIGNAL == SIGNAL_DIFF ? cb.diffMaxAccumulatedFrameNum : cb.specMaxAccumulatedFrameNum)));

    // ---- diffuse ----
    {
        float diffMaxAccumulatedFrameNum = cb.diffMaxAccumulatedFrameNum, diffMaxFastAccumulatedFrameNum = cb.diffMaxFastAccumulatedFrameNum;
        if (OPT && cb.hasHistoryConfidence) {  // TA:599-604
            const float inDiffConfidence = saturate(p.diffConfDummy.sampleLinear(prevUVSMB));
            diffMaxAccumulatedFrameNum *= inDiffConfidence;
            diffMaxFastAccumulatedFrameNum *= inDiffConfidence;
        }
        float diffuseAlpha = SMBReprojectionFound > 0.0f ? fmaxf(1.0f / (diffMaxAccumulatedFrameNum + 1.0f), 1.0f / historyLength) : 1.0f;
        float diffuseAlphaResponsive = SMBReprojectionFound > 0.0f ? fmaxf(1.0f / (diffMaxFastAccumulatedFrameNum + 1.0f), 1.0f / historyLength) : 1.0f;
        if (OPT && cb.diffCheckerboard != 2u && checkerboard != cb.diffCheckerboard && historyLength > 1.0f) {  // TA:611-620
            diffuseAlpha *= 1.0f - cb.checkerboardResolveAccumSpeed;
            diffuseAlphaResponsive *= 1.0f - cb.checkerboardResolveAccumSpeed;
        }
        p.outDiff.store(px, py, lerp(prevDiffuseSMB, f4(diffuseIllumination, diffuse2ndMoment), diffuseAlpha));
        p.outDiffFast.store(px, py, f4(lerp(prevDiffuseSMBResponsive, diffuseIllumination, diffuseAlphaResponsive), 0.0f));
        storeSh<SH>(p.outDiffSh, px, py, lerp(prevDiffuseSH, diffuseSH, diffuseAlpha));
        storeSh<SH>(p.outDiffShFast, px, py, lerp(prevDiffuseResponsiveSH, diffuseSH, diffuseAlphaResponsive));
    }
    p.outHistoryLength.store(px, py, historyLength / 255.0f);

    // ---- specular ----
    float specMaxAccumulatedFrameNum = cb.specMaxAccumulatedFrameNum, specMaxFastAccumulatedFrameNum = cb.specMaxFastAccumulatedFrameNum;
    if (OPT && cb.hasHistoryConfidence) {  // TA:642-647
        const float inSpecConfidence = saturate(p.specConfDummy.sampleLinear(prevUVSMB));
        specMaxAccumulatedFrameNum *= inSpecConfidence;
        specMaxFastAccumulatedFrameNum *= inSpecConfidence;
    }
    const float specHistoryFrames = fminf(specMaxAccumulatedFrameNum, historyLength);
    const float specHistoryResponsiveFrames = fminf(specMaxFastAccumulatedFrameNum, historyLength);
    const float hitDist = minHitDist3x3 == NRD_INF ? 0.0f : minHitDist3x3;

    float curvature = 0.0f;
    {
        const float2 uvForZeroParallax = cb.orthoMode == 0.0f ? prevUVSMB : pixelUv;
        float2 deltaUv = uvForZeroParallax - screenUv(cb.worldToClipPrev, prevWorldPos + cameraDelta);
        deltaUv = deltaUv * rectSize;
        deltaUv = deltaUv / fmaxf(smbParallaxInPixels1, 1.0f / 256.0f);

        auto edgePoint = [&](float2 d) -> float3 {
            const float3 x = currentWorldPosClip(cb, (pixelUv + d * rectSizeInv) * 2.0f - 1.0f, 1.0f);
            const float3 v = cb.orthoMode == 0.0f ? normalize(-x) : fwd;
            const float3 o = cb.orthoMode == 0.0f ? f3(0.0f) : x;
            return o + v * dot(currentWorldPos - o, currentNormal) / dot(currentNormal, v);  // line-plane intersection
        };
        const float3 x10 = edgePoint(make_float2(1.0f, 0.0f)), x01 = edgePoint(make_float2(0.0f, 1.0f));
        const float3 n10 = xyz(preload(px + 1, py)), n01 = xyz(preload(px, py + 1));

        float2 w = fabs2(deltaUv) + 1.0f / 256.0f;
        w = w / (w.x + w.y);
        float3 x = x10 * w.x + x01 * w.y;
        float3 n = normalize(n10 * w.x + n01 * w.y);

        const float dither = bayer4x4((uint32_t)px, (uint32_t)py, cb.frameIndex);
        const float edgeFix = 1.0f - pow5(NoV);
        float deltaUvLenFixed = smbParallaxInPixelsMin;
        deltaUvLenFixed *= 1.0f + edgeFix * (1.0f + cb.framerateScale * dither);
        float2 motionUvHigh = pixelUv + deltaUvLenFixed * deltaUv * rectSizeInv;
        motionUvHigh = (floor2(motionUvHigh * rectSize) + 0.5f) * rectSizeInv;
        if (deltaUvLenFixed > 1.0f && isInScreenNearest(motionUvHigh)) {
            const float2 uvScaled = clampUvToViewport(cb, motionUvHigh);
            const float zHigh = relaxViewZ(cb, p.viewZ.sampleNearest(uvScaled));
            const float3 xHigh = currentWorldPosClip(cb, motionUvHigh * 2.0f - 1.0f, zHigh);
            const float3 nHigh = xyz(unpackNormalRoughness(p.normalRoughness.sampleNearestRaw(uvScaled)));
            const float2 gp = geometryWeightParams(0.04f, minRectDim * pixelSize, currentWorldPos, currentNormal);
            float wg = nonExponentialWeight(dot(currentNormal, xHigh), gp.x, gp.y);
            wg = !relaxInRange(cb, zHigh) ? 0.0f : wg;
            const bool cmp = wg > 0.5f;
            n = cmp ? nHigh : n;
            x = cmp ? xHigh : x;
        }
        const float3 edge = x - currentWorldPos;
        curvature = dot(n - currentNormal, edge) / dot(edge, edge);
        if (curvature < 0.0f) {
            const float2 uv1 = screenUv(cb.worldToClipPrev, getXvirtual(hitDist, curvature, currentWorldPos, currentWorldPos, currentNormal, V, currentRoughness));
            const float2 uv2 = screenUv(cb.worldToClipPrev, currentWorldPos);
            const float a = length((uv1 - uv2) * rectSize);
            curvature *= a < 5.0f * smbParallaxInPixelsMax + rectSizeInv.x ? 1.0f : 0.0f;
        }
    }

    const float3 virtualWorldPos = getXvirtual(hitDist, curvature, currentWorldPos, prevWorldPos, currentNormal, V, currentRoughness);

    // ---- virtual-motion based history (TA:239-357) ----
    float4 prevSpecularVMB = f4(0.0f), prevSpecularVMBResponsive = f4(0.0f);
    float3 prevNormalVMB = currentNormal, prevSpecularVMBSH = f3(0.0f), prevSpecularVMBResponsiveSH = f3(0.0f);
    float prevRoughnessVMB = 0.0f, prevReflectionHitTVMB = cb.denoisingRange, VMBReprojectionFound;
    float2 prevUVVMB;
    {
        const float4 clip = mulM4(cb.worldToClipPrev, f4(virtualWorldPos, 1.0f));
        prevUVVMB = make_float2(clip.x / clip.w, clip.y / clip.w) * make_float2(0.5f, -0.5f) + 0.5f;
        prevUVVMB = currentMaterialID == cb.cameraAttachedReflectionMaterialID ? prevUVSMB : prevUVVMB;
        const float2 prevVirtualPixelPosFloat = prevUVVMB * rectSizePrev;
        const float2 fl = floor2(prevVirtualPixelPosFloat - 0.5f);
        const int ox = (int)fl.x, oy = (int)fl.y;
        Bilinear bilinear;
        bilinear.origin = fl;
        bilinear.weights = frac2(prevVirtualPixelPosFloat - 0.5f);

        const float3 cwp = currentWorldPos - cameraDelta;
        float4 thr = f4(disocclusionThreshold * (cb.orthoMode == 0.0f ? currentLinearZ : 1.0f));
        thr = thr * isInScreenBilinear(fl, rectSizePrev);
        thr = thr - NRD_EPS;

        const float4 zr = gatherR32(p.prevViewZ, ox, oy);
        const float4 prevMaterialIDs = gatherR8(p.prevMaterialID, ox, oy) * 255.0f;
        auto tapValid = [&](int dx, int dy, float zraw, float t, float mat) -> float {
            const float3 q = previousWorldPosPixel(cb, ox + dx, oy + dy, relaxViewZ(cb, zraw));
            const float v = fabsf(dot(cwp - q, currentNormal)) > t ? 0.0f : 1.0f;
            return v * (compareMaterials(currentMaterialID, mat, cb.specMinMaterial) ? 1.0f : 0.0f);
        };
        const float4 tapsValid = make_float4(tapValid(0, 0, zr.x, thr.x, prevMaterialIDs.x), tapValid(1, 0, zr.y, thr.y, prevMaterialIDs.y), tapValid(0, 1, zr.z, thr.z, prevMaterialIDs.z),
                                             tapValid(1, 1, zr.w, thr.w, prevMaterialIDs.w));
        const bool anyValid = tapsValid.x != 0.0f || tapsValid.y != 0.0f || tapsValid.z != 0.0f || tapsValid.w != 0.0f;
        const bool allValid = tapsValid.x != 0.0f && tapsValid.y != 0.0f && tapsValid.z != 0.0f && tapsValid.w != 0.0f;
        // The virtual-motion history is fetched with the validity gathers still in flight instead of behind their outcome ( the shader's `if( any valid )` ): an
        // invalid footprint takes the defaults below, as before. With the history filter requesting its 12 texels ahead of its bicubic / bilinear choice
        // ( HF_FETCH_ALL ) one memory round trip leaves the dependent chain: 556 -> 543 us per 1440p frame, results unchanged.
#ifndef RELAX_TA_VMB_COND
        constexpr bool kVmbAlways = true;
#else
        constexpr bool kVmbAlways = false;
#endif
        if (kVmbAlways || anyValid) {
            const float4 bilinearCustomW = bilinearCustomWeights(bilinear, tapsValid);
            const bool useBicubic = SMBReprojectionFound == 2.0f && allValid;
            const HistoryFilter hf(prevVirtualPixelPosFloat, resourceSizeInvPrev, bilinearCustomW, useBicubic);
            prevSpecularVMB = max4v(hf.color(p.historySpec), 0.0f);
            prevSpecularVMBResponsive = max4v(hf.color(p.historySpecFast), 0.0f);
            prevSpecularVMBSH = customWeightsSH<SH>(p.historySpecSh, ox, oy, bilinearCustomW);
            prevSpecularVMBResponsiveSH = customWeightsSH<SH>(p.historySpecShFast, ox, oy, bilinearCustomW);
            prevReflectionHitTVMB = fmaxf(0.001f, p.prevSpecHitDist.sampleLinear(prevUVVMB * resolutionScalePrev));
            const float4 prevNR = unpackPrevNormalRoughness(p.prevNormalRoughness.sampleLinear(prevUVVMB * resolutionScalePrev));
            prevNormalVMB = xyz(prevNR);
            prevRoughnessVMB = prevNR.w;
        }
        if (kVmbAlways && !anyValid) {
            prevSpecularVMB = prevSpecularVMBResponsive = f4(0.0f);
            prevSpecularVMBSH = prevSpecularVMBResponsiveSH = f3(0.0f);
            prevNormalVMB = currentNormal;
            prevRoughnessVMB = 0.0f;
            prevReflectionHitTVMB = cb.denoisingRange;
        }
        VMBReprojectionFound = allValid ? 1.0f : 0.0f;
    }

    const float4 D = specularDominantDirectionG2(currentNormal, V, currentRoughnessModified);
    float virtualHistoryAmount = VMBReprojectionFound * D.w;
    virtualHistoryAmount *= cb.orthoMode == 0.0f ? 1.0f : 0.75f;
    virtualHistoryAmount *= dot(prevNormalVMB, currentNormalAveraged) > 0.0f ? 1.0f : 0.0f;

    float2 uvDiff = prevUVVMB - prevUVSMB;
    const float uvDiffLengthInPixels = length(uvDiff * rectSize);
    float tanCurvature = fabsf(curvature * pixelSize);
    tanCurvature *= fmaxf(uvDiffLengthInPixels / fmaxf(NoV, 0.01f), 1.0f);
    const float curvatureAngle = atanf(tanCurvature);

    const float lobeHalfAngle = fmaxf(atanf(relaxSpecLobeTanHalfAngle(currentRoughnessModified)), RELAX_NORMAL_ULP);
    const float normalWeight = encodingAwareNormalWeight(currentNormal, prevNormalVMB, lobeHalfAngle, curvatureAngle, RELAX_NORMAL_ULP);
    virtualHistoryAmount *= lerp(1.0f - saturate(uvDiffLengthInPixels), 1.0f, normalWeight);

    const float2 relaxedRoughnessP = relaxedRoughnessWeightParams(currentRoughness * currentRoughness, cb.roughnessFraction);
    float virtualRoughnessWeight = computeWeight(prevRoughnessVMB * prevRoughnessVMB, relaxedRoughnessP.x, relaxedRoughnessP.y);
    virtualRoughnessWeight = lerp(1.0f - saturate(uvDiffLengthInPixels), 1.0f, virtualRoughnessWeight);
    virtualHistoryAmount *= cb.orthoMode == 0.0f ? virtualRoughnessWeight : 1.0f;
    float specVMBConfidence = virtualRoughnessWeight * 0.9f + 0.1f;

    uvDiff = uvDiff * rsqrtSafe(dot(uvDiff, uvDiff));
    uvDiff = uvDiff / rectSizePrev;
    uvDiff = uvDiff * (saturate(uvDiffLengthInPixels / 0.1f) + uvDiffLengthInPixels / 2.0f);
    const float2 backUV1 = prevUVVMB + 1.0f * uvDiff, backUV2 = prevUVVMB + 2.0f * uvDiff;
    const float4 backNR1 = unpackPrevNormalRoughness(p.prevNormalRoughness.sampleLinear(backUV1 * resolutionScalePrev));
    const float4 backNR2 = unpackPrevNormalRoughness(p.prevNormalRoughness.sampleLinear(backUV2 * resolutionScalePrev));
    float prevPrevNormalWeight = isInScreenNearest(backUV1) ? encodingAwareNormalWeight(prevNormalVMB, xyz(backNR1), lobeHalfAngle, curvatureAngle * 2.0f, RELAX_NORMAL_ULP) : 1.0f;
    prevPrevNormalWeight *= isInScreenNearest(backUV2) ? encodingAwareNormalWeight(prevNormalVMB, xyz(backNR2), lobeHalfAngle, curvatureAngle * 3.0f, RELAX_NORMAL_ULP) : 1.0f;
    virtualHistoryAmount *= 0.33f + 0.67f * prevPrevNormalWeight;
    specVMBConfidence *= 0.33f + 0.67f * prevPrevNormalWeight;
    float rw = computeWeight(backNR1.w * backNR1.w, relaxedRoughnessP.x, relaxedRoughnessP.y);
    rw *= computeWeight(backNR2.w * backNR2.w, relaxedRoughnessP.x, relaxedRoughnessP.y);
    virtualHistoryAmount *= cb.orthoMode == 0.0f ? rw * 0.9f + 0.1f : 1.0f;

    const float SMC = specMagicCurve(currentRoughnessModified);
    const float hitDistC = lerp(specularIllumination.w, prevReflectionHitTSMB, SMC);
    const float hitDist1 = thinLens(hitDistC, curvature), hitDist2 = thinLens(prevReflectionHitTVMB, curvature);
    const float maxDist = fmaxf(hitDist1, hitDist2);
    const float dHitT = fabsf(hitDist1 - hitDist2);
    const float dHitTMultiplier = lerp(20.0f, 0.0f, SMC);
    float virtualHistoryHitDistConfidence = 1.0f - saturate(dHitTMultiplier * dHitT / (currentLinearZ + maxDist));
    virtualHistoryHitDistConfidence = lerp(virtualHistoryHitDistConfidence, 1.0f, SMC);

    const float virtualWorldPosLength = length(virtualWorldPos);
    const float hitDistForTrackingPrev = prevSpecularVMBResponsive.w;
    const float3 prevVirtualWorldPos = getXvirtual(hitDistForTrackingPrev, curvature, currentWorldPos, prevWorldPos, currentNormal, V, currentRoughness);
    const float virtualWorldPosLengthPrev = length(prevVirtualWorldPos);
    float2 prevUVVMBTest = screenUv(cb.worldToClipPrev, prevVirtualWorldPos);
    prevUVVMBTest = currentMaterialID == cb.cameraAttachedReflectionMaterialID ? prevUVSMB : prevUVVMBTest;
    const float lobeTanHalfAngle = fmaxf(relaxSpecLobeTanHalfAngle(currentRoughness, 0.6f), 0.5f * rectSizeInv.x);
    const float unproj1 = fminf(hitDist, hitDistForTrackingPrev) / pixelRadiusToWorld(cb.unproject, cb.orthoMode, 1.0f, fmaxf(virtualWorldPosLength, virtualWorldPosLengthPrev));
    const float lobeRadiusInPixels = lobeTanHalfAngle * unproj1;
    const float deltaParallaxInPixels = length((prevUVVMBTest - prevUVVMB) * rectSize);
    virtualHistoryHitDistConfidence *= smoothStep(lobeRadiusInPixels + 0.25f, 0.0f, deltaParallaxInPixels);

    const float specSMBConfidence = (SMBReprojectionFound > 0.0f ? 1.0f : 0.0f) * encodingAwareNormalWeight(V, Vprev, lobeHalfAngle * NoV / cb.framerateScale, 0.0f, 0.0f);
    float specSMBAlpha = 1.0f - specSMBConfidence;
    specSMBAlpha = fmaxf(specSMBAlpha, 1.0f / (1.0f + specHistoryFrames));
    float specSMBResponsiveAlpha = fmaxf(specSMBAlpha, 1.0f / (1.0f + specHistoryResponsiveFrames));
    // pixels the checkerboard skipped this frame accumulate faster while the camera is (nearly) static ( TA:866-875, 893-899 )
    const bool specResolved = OPT && cb.specCheckerboard != 2u && checkerboard != cb.specCheckerboard && smbParallaxInPixelsMax < 0.5f;
    if (specResolved) {
        const float k = 1.0f - cb.checkerboardResolveAccumSpeed * (SMBReprojectionFound > 0.0f ? 1.0f : 0.0f);
        specSMBAlpha *= k;
        specSMBResponsiveAlpha *= k;
    }

    const float3 accSMBrgb = lerp(xyz(prevSpecularSMB), xyz(specularIllumination), specSMBAlpha);
    const float accSMBw = lerp(prevReflectionHitTSMB, specularIllumination.w, fmaxf(specSMBAlpha, 0.1f));
    const float accM2SMB = lerp(prevSpecularSMB.w, specular2ndMoment, specSMBAlpha);
    const float3 accSMBResponsive = lerp(prevSpecularSMBResponsive, xyz(specularIllumination), specSMBResponsiveAlpha);

    float specVMBAlpha = 1.0f - specVMBConfidence;
    float specVMBResponsiveAlpha = 1.0f - specVMBConfidence * virtualHistoryHitDistConfidence;
    float specVMBHitTAlpha = specVMBResponsiveAlpha;
    specVMBAlpha = fmaxf(specVMBAlpha, 1.0f / (1.0f + specHistoryFrames));
    specVMBResponsiveAlpha = fmaxf(specVMBResponsiveAlpha, 1.0f / (1.0f + specHistoryResponsiveFrames));
    specVMBHitTAlpha = fmaxf(specVMBHitTAlpha, 1.0f / (1.0f + specHistoryFrames));
    if (specResolved) {
        const float k = 1.0f - cb.checkerboardResolveAccumSpeed * (VMBReprojectionFound > 0.0f ? 1.0f : 0.0f);
        specVMBAlpha *= k;
        specVMBResponsiveAlpha *= k;
        specVMBHitTAlpha *= k;
    }

    const float3 accVMBrgb = lerp(xyz(prevSpecularVMB), xyz(specularIllumination), specVMBAlpha);
    const float accVMBw = lerp(prevReflectionHitTVMB, specularIllumination.w, fmaxf(specVMBHitTAlpha, 0.1f));
    const float accM2VMB = lerp(prevSpecularVMB.w, specular2ndMoment, specVMBAlpha);
    const float3 accVMBResponsive = lerp(xyz(prevSpecularVMBResponsive), xyz(specularIllumination), specVMBResponsiveAlpha);

    virtualHistoryAmount *= saturate(specVMBConfidence / (specSMBConfidence + NRD_EPS));

    const float accumulatedReflectionHitT = lerp(accSMBw, accVMBw, virtualHistoryAmount);
    const float3 accumulatedSpecular = lerp(accSMBrgb, accVMBrgb, virtualHistoryAmount);
    const float3 accumulatedSpecularResponsive = lerp(accSMBResponsive, accVMBResponsive, virtualHistoryAmount);
    float accumulatedSpecular2ndMoment = lerp(accM2SMB, accM2VMB, virtualHistoryAmount);

    const float3 accSMBSH = lerp(prevSpecularSMBSH, specularSH, specSMBAlpha), accSMBRespSH = lerp(prevSpecularSMBResponsiveSH, specularSH, specSMBResponsiveAlpha);
    const float3 accVMBSH = lerp(prevSpecularVMBSH, specularSH, specVMBAlpha), accVMBRespSH = lerp(prevSpecularVMBResponsiveSH, specularSH, specVMBResponsiveAlpha);
    storeSh<SH>(p.outSpecSh, px, py, lerp(accSMBSH, accVMBSH, virtualHistoryAmount));
    storeSh<SH>(p.outSpecShFast, px, py, lerp(accSMBRespSH, accVMBRespSH, virtualHistoryAmount));

    const float specularHistoryConfidence = lerp(specSMBConfidence, specVMBConfidence, virtualHistoryAmount);
    if (accumulatedSpecular2ndMoment == 0.0f) accumulatedSpecular2ndMoment = cb.specVarianceBoost * (1.0f - specularHistoryConfidence);

    p.outSpec.store(px, py, f4(accumulatedSpecular, accumulatedSpecular2ndMoment));
    p.outSpecFast.store(px, py, f4(accumulatedSpecularResponsive, hitDist));
    p.outSpecHitDist.store(px, py, accumulatedReflectionHitT);
    p.outSpecReprojectionConfidence.store(px, py, specularHistoryConfidence);
}

// ---------------------------------------------------------------------------------------------------------------
template <bool SH, int SIGNAL>
__global__ void __launch_bounds__(BLOCK_W* BLOCK_H) relaxHistoryFixKernel(const __grid_constant__ RelaxConstants cb, const __grid_constant__ RelaxHistoryFixParamsT<SIGNAL> p, int ctaY0) {
    pdlEntry();
    const int px = blockIdx.x * BLOCK_W + threadIdx.x, py = (blockIdx.y + ctaY0) * BLOCK_H + threadIdx.y;
    if (p.tiles.load(px >> 4, py >> 4) != 0.0f || px >= cb.rectSize[0] || py >= cb.rectSize[1]) return;
    const float centerViewZ = relaxViewZ(cb, p.viewZ.load(px, py));
    const float historyLength = 255.0f * p.historyLength.load(px, py);
    if (!relaxInRange(cb, centerViewZ) || historyLength > cb.historyFixFrameNum || cb.historyFixFrameNum == 1.0f) return;

    float centerMaterialID;
    const float4 centerNormalRoughness = unpackNormalRoughness(p.normalRoughness.loadRaw(px, py), centerMaterialID);
    const float3 centerNormal = xyz(centerNormalRoughness);
    const float3 centerWorldPos = currentWorldPosPixel(cb, px, py, centerViewZ);
    const float3 centerV = -normalize(centerWorldPos);
    const float depthThreshold = cb.depthThreshold * (cb.orthoMode == 0.0f ? centerViewZ : 1.0f);
    const float2 rectSize = make_float2((float)cb.rectSize[0], (float)cb.rectSize[1]), rectSizeInv = make_float2(cb.rectSizeInv[0], cb.rectSizeInv[1]);

    float4 diffuseSum = p.diff.load(px, py), specularSum = p.spec.load(px, py);
    float3 diffuseSumSH = loadSh<SH>(p.diffSh, px, py), specularSumSH = loadSh<SH>(p.specSh, px, py);
    float diffuseWSum = 1.0f, specularWSum = 1.0f;
    const float2 specularNormalWeightP = normalWeightParamsAtrous(centerNormalRoughness.w, 5.0f, 1.0f, 0.0f, cb.lobeAngleFraction, cb.specLobeAngleSlack);
    const float normalPower = fmaxf(cb.historyFixEdgeStoppingNormalPower, 0.01f);

    const float baseStride = centerMaterialID == cb.historyFixAlternatePixelStrideMaterialID ? cb.historyFixAlternatePixelStride : cb.historyFixBasePixelStride;
    const float r = roundNe(baseStride / (1.0f + historyLength));
    for (int j = -2; j <= 2; j++)
        for (int i = -2; i <= 2; i++) {
            if (i == 0 && j == 0) continue;
            int sx = (int)((float)px + (float)i * r), sy = (int)((float)py + (float)j * r);
            const float2 uv = mirrorUv(make_float2((float)sx + 0.5f, (float)sy + 0.5f) * rectSizeInv);
            sx = (int)(uv.x * rectSize.x);
            sy = (int)(uv.y * rectSize.y);

            float sampleMaterialID;
            const float3 sampleNormal = xyz(unpackNormalRoughness(p.normalRoughness.loadRaw(sx, sy), sampleMaterialID));
            const float sampleViewZ = relaxViewZ(cb, p.viewZ.load(sx, sy));
            const float3 sampleWorldPos = currentWorldPosPixel(cb, sx, sy, sampleViewZ);
            float geometryWeight = planeDistanceWeightAtrous(centerWorldPos, centerNormal, sampleWorldPos, depthThreshold);
            geometryWeight = relaxInRange(cb, sampleViewZ) ? geometryWeight : 0.0f;

            float diffuseW = geometryWeight;
            diffuseW *= powf(fmaxf(0.01f, dot(centerNormal, sampleNormal)), normalPower);
            diffuseW *= compareMaterials(sampleMaterialID, centerMaterialID, cb.diffMinMaterial) ? 1.0f : 0.0f;
            if (diffuseW > 1e-4f) {
                diffuseSum += p.diff.load(sx, sy) * diffuseW;
                diffuseSumSH += loadSh<SH>(p.diffSh, sx, sy) * diffuseW;
                diffuseWSum += diffuseW;
            }
            const float3 sampleV = -normalize(sampleWorldPos + cb.roughnessEdgeStoppingRelaxation * centerWorldPos);
            float specularW = geometryWeight;
            specularW *= specularNormalWeightAtrous(specularNormalWeightP, centerNormal, sampleNormal, centerV, sampleV);
            specularW *= compareMaterials(sampleMaterialID, centerMaterialID, cb.specMinMaterial) ? 1.0f : 0.0f;
            if (specularW > 1e-4f) {
                specularSum += p.spec.load(sx, sy) * specularW;
                specularSumSH += loadSh<SH>(p.specSh, sx, sy) * specularW;
                specularWSum += specularW;
            }
        }
    p.outDiff.store(px, py, diffuseSum / diffuseWSum);
    storeSh<SH>(p.outDiffSh, px, py, diffuseSumSH / diffuseWSum);
    p.outSpec.store(px, py, specularSum / specularWSum);
    storeSh<SH>(p.outSpecSh, px, py, specularSumSH / specularWSum);
}

// ---------------------------------------------------------------------------------------------------------------
// History clamping: the 5x5 neighbourhood of { responsive history in YCoCg, validity } and { noisy input, luminance^2 } lives in shared
// memory (36x12 texels per 32x8 CTA), converted once per texel instead of once per tap.
// The pass is bound by shared-memory bandwidth, not by arithmetic or DRAM: 25 taps x 2 lobes x 2 LDS.128 per pixel are ~160 us of LDS wavefronts per 1440p frame on
// their own. So a thread owns HC_ROWS vertically adjacent pixels: their 5x5 windows overlap in 5 x ( HC_ROWS + 4 ) texels, every staged texel is read once per
// thread and added to the sums of each pixel whose window holds it ( 30 tap reads instead of 50 for two pixels ). Each pixel still adds its 25 taps in the
// shader's order ( dx outer, dy inner ), so its sums — and everything after them — are bit for bit what one thread per pixel computed.
constexpr int HC_BORDER = 2, HC_ROWS = 2, HC_PIXELS_H = BLOCK_H * HC_ROWS, HC_TILE_W = BLOCK_W + 2 * HC_BORDER, HC_TILE_H = HC_PIXELS_H + 2 * HC_BORDER;

// One lobe (the specular and diffuse halves of the shader differ in three constants only) for the HC_ROWS pixels ( px, py0 + r ) of a thread; active[ r ]: the
// pixel is processed ( inside the rect / strip, not sky, in range )
template <bool SPEC, bool SH, class TEX>
NRD_DEV void historyClampingLobe(const RelaxConstants& cb, const float4 (*sFast)[HC_TILE_W], const float4 (*sNoisy)[HC_TILE_W], int px, int py0, const bool* active, const float* historyLengths,
                                 const TEX& slowTex, const TEX& fastTex, const TEX& shTex, const TEX& shFastTex, const TEX& outSlow,
                                 const TEX& outFast, const TEX& outSh, const TEX& outShFast, float maxFast, float maxSlow) {
    const int sx = threadIdx.x + HC_BORDER, sy0 = threadIdx.y * HC_ROWS + HC_BORDER;
    float3 m1s[HC_ROWS], m2s[HC_ROWS], noisyM1s[HC_ROWS];
    float noisyM2s[HC_ROWS], sums[HC_ROWS];
#pragma unroll
    for (int r = 0; r < HC_ROWS; r++) {
        m1s[r] = m2s[r] = noisyM1s[r] = f3(0.0f);
        noisyM2s[r] = sums[r] = 0.0f;
    }
#pragma unroll
    for (int dx = -2; dx <= 2; dx++)
#pragma unroll
        for (int row = -2; row <= HC_ROWS + 1; row++) {   // tile row sy0 + row: tap dy = row - r of pixel r
            const float4 f = sFast[sy0 + row][sx + dx];
            const float4 n = sNoisy[sy0 + row][sx + dx];
            if (f.w != 0.0f) {
                const float3 c = xyz(f), c2 = c * c;
#pragma unroll
                for (int r = 0; r < HC_ROWS; r++)
                    if (row - r >= -2 && row - r <= 2) {
                        m1s[r] += c;
                        m2s[r] += c2;
                        noisyM1s[r] += xyz(n);
                        noisyM2s[r] += n.w;
                        sums[r] += 1.0f;
                    }
            }
        }
#pragma unroll
    for (int r = 0; r < HC_ROWS; r++) {
    if (!active[r]) continue;
    const int py = py0 + r, sy = sy0 + r;
    const float historyLength = historyLengths[r];
    float3 m1 = m1s[r], m2 = m2s[r], noisyM1 = noisyM1s[r];
    float noisyM2 = noisyM2s[r];
    const float sum = sums[r];
    m1 = m1 / sum;
    m2 = m2 / sum;
    noisyM1 = noisyM1 / sum;
    noisyM2 /= sum;
    const float3 sigma = sqrt3v(max3(f3(0.0f), m2 - m1 * m1));
    float3 colorMin = m1 - cb.fastHistoryClampingSigmaScale * sigma, colorMax = m1 + cb.fastHistoryClampingSigmaScale * sigma;

    const float3 responsiveCenterYCoCg = xyz(sFast[sy][sx]);
    const float responsiveCenterAlpha = SPEC ? fastTex.load(px, py).w : 0.0f;
    colorMin = min3v(colorMin, responsiveCenterYCoCg);
    colorMax = max3(colorMax, responsiveCenterYCoCg);

    const float4 slow = slowTex.load(px, py);
    const float3 slowYCoCg = rgbToYCoCg(xyz(slow));
    float3 clampedYCoCg = slowYCoCg;
    if (maxFast < maxSlow) clampedYCoCg = clamp3v(slowYCoCg, colorMin, colorMax);
    const float3 clamped = yCoCgToRgb(clampedYCoCg);

    float4 outSlowV = f4(clamped, slow.w);
    const float3 responsiveCenter = yCoCgToRgb(responsiveCenterYCoCg);
    float4 outFastV = f4(responsiveCenter, responsiveCenterAlpha);
    const bool fixed = historyLength <= cb.historyFixFrameNum;
    if (fixed) outSlowV = SPEC ? outFastV : f4(xyz(outFastV), outSlowV.w);

    float clampingFactor = (clampedYCoCg.x - slowYCoCg.x) == 0.0f ? 0.0f : saturate((clampedYCoCg.x - slowYCoCg.x) / (responsiveCenterYCoCg.x - slowYCoCg.x));
    if (fixed) clampingFactor = 1.0f;

    float historyDifferenceL = (SPEC ? 0.33f : 1.0f) * RELAX_ANTILAG_ACCELERATION_AMOUNT_SCALE * cb.historyAccelerationAmount * luminance(fabs3(responsiveCenter - xyz(slow)));
    historyDifferenceL *= clampingFactor;
    if (fixed) historyDifferenceL = 0.0f;

    const float3 distanceToNoisy = noisyM1 - responsiveCenter;
    const float distanceToNoisyL = luminance(fabs3(distanceToNoisy));
    float3 acceleration = distanceToNoisyL == 0.0f ? f3(0.0f) : distanceToNoisy * historyDifferenceL / distanceToNoisyL;
    const float accelerationL = luminance(fabs3(acceleration));
    const float accelerationRatio = accelerationL == 0.0f ? 0.0f : distanceToNoisyL / accelerationL;
    if (accelerationRatio < 1.0f) acceleration = acceleration * accelerationRatio;
    if (accelerationRatio <= 0.0f) acceleration = f3(0.0f);
    outSlowV = f4(xyz(outSlowV) + acceleration, outSlowV.w);
    outFastV = f4(xyz(outFastV) + acceleration, outFastV.w);

    const float slowL = luminance(xyz(slow));
    const float noisyInputL = luminance(noisyM1);
    const float temporalSigma = cb.historyResetTemporalSigmaScale * sqrtf(fmaxf(0.0f, noisyM2 - noisyInputL * noisyInputL));
    const float spatialSigma = cb.historyResetSpatialSigmaScale * sigma.x;
    float resetAmount = (SPEC ? 0.5f : 1.0f) * cb.historyResetAmount * fmaxf(0.0f, fabsf(slowL - noisyInputL) - spatialSigma - temporalSigma) /
                        (1.0e-6f + fmaxf(slowL, noisyInputL) + spatialSigma + temporalSigma);
    resetAmount = saturate(resetAmount);
    const float3 noisyCenter = xyz(sNoisy[sy][sx]);
    outSlowV = f4(lerp(xyz(outSlowV), noisyCenter, resetAmount), outSlowV.w);
    outFastV = f4(lerp(xyz(outFastV), noisyCenter, resetAmount), outFastV.w);

    const float outL = luminance(xyz(outSlowV));
    outSlowV.w = fmaxf(0.0f, outSlowV.w + (outL * outL - slowL * slowL));

    outSlow.store(px, py, outSlowV);
    outFast.store(px, py, outFastV);
    const float3 sh = loadSh<SH>(shTex, px, py), shFast = loadSh<SH>(shFastTex, px, py);
    storeSh<SH>(outSh, px, py, lerp(sh, shFast, clampingFactor));
    storeSh<SH>(outShFast, px, py, shFast);
    }
}

// ctaY0: first CTA row in units of HC_PIXELS_H ( = 16 ) pixel rows; rowEnd: end of the strip ( or of the rect )
template <bool SH, int SIGNAL>
__global__ void __launch_bounds__(BLOCK_W* BLOCK_H) relaxHistoryClampingKernel(const __grid_constant__ RelaxConstants cb, const __grid_constant__ RelaxHistoryClampingParamsT<SIGNAL> p, int ctaY0, int rowEnd) {
    pdlEntry();
    __shared__ float4 sSpecFast[HC_TILE_H][HC_TILE_W], sSpecNoisy[HC_TILE_H][HC_TILE_W], sDiffFast[HC_TILE_H][HC_TILE_W], sDiffNoisy[HC_TILE_H][HC_TILE_W];
    const int px = blockIdx.x * BLOCK_W + threadIdx.x, py0 = (blockIdx.y + ctaY0) * HC_PIXELS_H + threadIdx.y * HC_ROWS;
    // the CTA covers two 16x16 tiles of one tile row
    const float skyL = p.tiles.load((blockIdx.x * BLOCK_W) >> 4, py0 >> 4), skyR = p.tiles.load((blockIdx.x * BLOCK_W + 16) >> 4, py0 >> 4);
    if (skyL != 0.0f && skyR != 0.0f) return;
    {
        const int baseX = blockIdx.x * BLOCK_W - HC_BORDER, baseY = (blockIdx.y + ctaY0) * HC_PIXELS_H - HC_BORDER;
        const int maxX = cb.rectSize[0] - 1, maxY = cb.rectSize[1] - 1;
        for (int i = threadIdx.y * BLOCK_W + threadIdx.x; i < HC_TILE_W * HC_TILE_H; i += BLOCK_W * BLOCK_H) {
            const int tx = i % HC_TILE_W, ty = i / HC_TILE_W;
            const int gx = clampi(baseX + tx, 0, maxX), gy = clampi(baseY + ty, 0, maxY);
            const float valid = relaxInRange(cb, p.viewZ.load(gx, gy)) ? 1.0f : 0.0f;  // raw viewZ, as in the shader's Preload( )
            const float3 sn = xyz(p.specNoisy.load(gx, gy)), dn = xyz(p.diffNoisy.load(gx, gy));
            const float sl = luminance(sn), dl = luminance(dn);
            sSpecFast[ty][tx] = f4(rgbToYCoCg(xyz(p.specFast.load(gx, gy))), valid);
            sSpecNoisy[ty][tx] = f4(sn, sl * sl);
            sDiffFast[ty][tx] = f4(rgbToYCoCg(xyz(p.diffFast.load(gx, gy))), valid);
            sDiffNoisy[ty][tx] = f4(dn, dl * dl);
        }
    }
    __syncthreads();
    const float isSky = threadIdx.x < 16 ? skyL : skyR;
    if (isSky != 0.0f || px >= cb.rectSize[0]) return;
    bool active[HC_ROWS];
    float historyLengths[HC_ROWS];
    bool any = false;
#pragma unroll
    for (int r = 0; r < HC_ROWS; r++) {
        // ( the validity flag of the staged centre texel: viewZ in the denoising range )
        active[r] = py0 + r < rowEnd && py0 + r < cb.rectSize[1] && sSpecFast[threadIdx.y * HC_ROWS + r + HC_BORDER][threadIdx.x + HC_BORDER].w != 0.0f;
        historyLengths[r] = active[r] ? 255.0f * p.historyLength.load(px, py0 + r) : 0.0f;
        any = any || active[r];
    }
    if (!any) return;
    historyClampingLobe<true, SH>(cb, sSpecFast, sSpecNoisy, px, py0, active, historyLengths, p.spec, p.specFast, p.specSh, p.specShFast, p.outSpec, p.outSpecFast, p.outSpecSh, p.outSpecShFast,
                              cb.specMaxFastAccumulatedFrameNum, cb.specMaxAccumulatedFrameNum);
    historyClampingLobe<false, SH>(cb, sDiffFast, sDiffNoisy, px, py0, active, historyLengths, p.diff, p.diffFast, p.diffSh, p.diffShFast, p.outDiff, p.outDiffFast, p.outDiffSh, p.outDiffShFast,
                               cb.diffMaxFastAccumulatedFrameNum, cb.diffMaxAccumulatedFrameNum);
#pragma unroll
    for (int r = 0; r < HC_ROWS; r++)
        if (active[r]) p.outHistoryLength.store(px, py0 + r, historyLengths[r] / 255.0f);
}

template <int SIGNAL>
__global__ void __launch_bounds__(BLOCK_W* BLOCK_H) relaxCopyKernel(const __grid_constant__ RelaxCopyParamsT<SIGNAL> p, int ctaY0) {
    pdlEntry();
    const int px = blockIdx.x * BLOCK_W + threadIdx.x, py = (blockIdx.y + ctaY0) * BLOCK_H + threadIdx.y;
    if constexpr ((SIGNAL & SIGNAL_SPEC) != 0)
        if (p.outSpec.inside(px, py)) *p.outSpec.template ptrw<uint2>(px, py) = p.spec.inside(px, py) ? p.spec.fetchRaw(px, py) : make_uint2(0u, 0u);
    if constexpr ((SIGNAL & SIGNAL_DIFF) != 0)
        if (p.outDiff.inside(px, py)) *p.outDiff.template ptrw<uint2>(px, py) = p.diff.inside(px, py) ? p.diff.fetchRaw(px, py) : make_uint2(0u, 0u);
}

// ---------------------------------------------------------------------------------------------------------------
template <class TEX> NRD_DEV float4 rcrs(const RelaxConstants& cb, const TexNR& normalRoughness, const TEX& tex, int px, int py, float minMaterial, float centerMaterialID) {
    const float4 center = tex.load(px, py);
    const float centerL = luminance(xyz(center));
    float maxL = -1.0f, minL = 1.0e6f;
    int maxXc = px, maxYc = py, minXc = px, minYc = py;
#pragma unroll
    for (int yy = -1; yy <= 1; yy++)
#pragma unroll
        for (int xx = -1; xx <= 1; xx++) {
            const int x = px + xx, y = py + yy;
            if ((xx == 0 && yy == 0) || x < 0 || y < 0 || x >= cb.rectSize[0] || y >= cb.rectSize[1]) continue;
            const float sampleL = luminance(xyz(tex.load(x, y)));
            if (compareMaterials(materialFromRaw(normalRoughness.loadRaw(x, y)), centerMaterialID, minMaterial)) {
                if (sampleL > maxL) { maxL = sampleL; maxXc = x; maxYc = y; }
                if (sampleL < minL) { minL = sampleL; minXc = x; minYc = y; }
            }
        }
    int sx = px, sy = py;
    if (centerL > maxL) { sx = maxXc; sy = maxYc; }
    if (centerL < minL) { sx = minXc; sy = minYc; }
    return f4(xyz(tex.load(sx, sy)), center.w);
}

template <int SIGNAL>
__global__ void __launch_bounds__(BLOCK_W* BLOCK_H) relaxAntiFireflyKernel(const __grid_constant__ RelaxConstants cb, const __grid_constant__ RelaxAntiFireflyParamsT<SIGNAL> p, int ctaY0) {
    pdlEntry();
    const int px = blockIdx.x * BLOCK_W + threadIdx.x, py = (blockIdx.y + ctaY0) * BLOCK_H + threadIdx.y;
    if (p.tiles.load(px >> 4, py >> 4) != 0.0f || px >= cb.rectSize[0] || py >= cb.rectSize[1]) return;
    if (!relaxInRange(cb, relaxViewZ(cb, p.viewZ.load(px, py)))) return;
    const float centerMaterialID = materialFromRaw(p.normalRoughness.loadRaw(px, py));
    p.outSpec.store(px, py, rcrs(cb, p.normalRoughness, p.spec, px, py, cb.specMinMaterial, centerMaterialID));
    p.outDiff.store(px, py, rcrs(cb, p.normalRoughness, p.diff, px, py, cb.diffMinMaterial, centerMaterialID));
}

// ---------------------------------------------------------------------------------------------------------------
// RELAX_HitDistReconstruction.cs.hlsl:21-158 ( 3x3 / 5x5 ): hit distances of pixels whose ray carried none ( hitDist = 0 ) are rebuilt from
// the neighbours on the same surface. The ( 32 + 2B ) x ( 8 + 2B ) neighbourhood of { normal } and { spec hitDist, diff hitDist, viewZ } is
// staged in shared memory, normals decoded once per texel. One permutation serves SH and RADIANCE ( only .w of the SH0 textures changes ).
struct RelaxHitDistReconstructionParams { TexR8 tiles; TexNR normalRoughness; TexR32F viewZ; TexRGBA16F spec, diff, outSpec, outDiff; };
template <int BORDER>
__global__ void __launch_bounds__(BLOCK_W* BLOCK_H) relaxHitDistReconstructionKernel(const __grid_constant__ RelaxConstants cb, const __grid_constant__ RelaxHitDistReconstructionParams p, int ctaY0) {
    pdlEntry();
    constexpr int TW = BLOCK_W + 2 * BORDER, TH = BLOCK_H + 2 * BORDER;
    __shared__ float3 sNormal[TH][TW];
    __shared__ float3 sHitDistViewZ[TH][TW];
    const int px = blockIdx.x * BLOCK_W + threadIdx.x, py = (blockIdx.y + ctaY0) * BLOCK_H + threadIdx.y;
    const float skyL = p.tiles.load((blockIdx.x * BLOCK_W) >> 4, py >> 4), skyR = p.tiles.load((blockIdx.x * BLOCK_W + 16) >> 4, py >> 4);
    if (skyL != 0.0f && skyR != 0.0f) return;
    {
        const int baseX = blockIdx.x * BLOCK_W - BORDER, baseY = (blockIdx.y + ctaY0) * BLOCK_H - BORDER;
        for (int i = threadIdx.y * BLOCK_W + threadIdx.x; i < TW * TH; i += BLOCK_W * BLOCK_H) {
            const int tx = i % TW, ty = i / TW;
            const int gx = clampi(baseX + tx, 0, cb.rectSize[0] - 1), gy = clampi(baseY + ty, 0, cb.rectSize[1] - 1);
            sNormal[ty][tx] = xyz(unpackNormalRoughness(p.normalRoughness.loadRaw(gx, gy)));
            sHitDistViewZ[ty][tx] = make_float3(p.spec.load(gx, gy).w, p.diff.load(gx, gy).w, relaxViewZ(cb, p.viewZ.load(gx, gy)));
        }
    }
    __syncthreads();
    const float isSky = threadIdx.x < 16 ? skyL : skyR;
    if (isSky != 0.0f || px >= cb.rectSize[0] || py >= cb.rectSize[1]) return;
    const float3 center = sHitDistViewZ[threadIdx.y + BORDER][threadIdx.x + BORDER];
    const float centerViewZ = center.z;
    if (!relaxInRange(cb, centerViewZ)) return;

    const float4 normalAndRoughness = unpackNormalRoughness(p.normalRoughness.loadRaw(px, py));
    const float3 centerNormal = xyz(normalAndRoughness);
    const float centerRoughness = normalAndRoughness.w;
    const float2 rectSizeInv = make_float2(cb.rectSizeInv[0], cb.rectSizeInv[1]);
    const float2 pixelUv = make_float2((float)px + 0.5f, (float)py + 0.5f) * rectSizeInv;
    const float2 relaxedRoughnessP = relaxedRoughnessWeightParams(centerRoughness * centerRoughness);
    const float specularNormalWeightParam = normalWeightParam(1.0f, 1.0f, centerRoughness), diffuseNormalWeightParam = normalWeightParam(1.0f, 1.0f);
    // the shader weights every tap by the CENTER roughness ( :121 ): a per-pixel constant
    const float roughnessW = exponentialWeight(centerRoughness * centerRoughness, relaxedRoughnessP.x, relaxedRoughnessP.y);

    float sumSpecularWeight = center.x != 0.0f ? 1000.0f : 0.0f, sumDiffuseWeight = center.y != 0.0f ? 1000.0f : 0.0f;
    float sumSpecularHitDist = center.x * sumSpecularWeight, sumDiffuseHitDist = center.y * sumDiffuseWeight;
#pragma unroll
    for (int j = 0; j <= BORDER * 2; j++)
#pragma unroll
        for (int i = 0; i <= BORDER * 2; i++) {
            if (i == BORDER && j == BORDER) continue;
            const float2 o = make_float2((float)(i - BORDER), (float)(j - BORDER));
            const float3 sampleNormal = sNormal[threadIdx.y + j][threadIdx.x + i];
            const float3 data = sHitDistViewZ[threadIdx.y + j][threadIdx.x + i];
            const float angle = acosApproxPositive(dot(centerNormal, sampleNormal));
            float w = isInScreenNearest(pixelUv + o * rectSizeInv) ? 1.0f : 0.0f;
            w *= relaxInRange(cb, data.z) ? 1.0f : 0.0f;
            w *= gaussianWeight(length(o) * 0.5f);
            w *= bilateralWeight(data.z, centerViewZ);

            float specularWeight = w * exponentialWeight(angle, specularNormalWeightParam, 0.0f);
            specularWeight *= roughnessW;
            const float sampleSpecularHitDist = specularWeight == 0.0f ? 0.0f : data.x;
            specularWeight *= sampleSpecularHitDist != 0.0f ? 1.0f : 0.0f;
            sumSpecularHitDist += sampleSpecularHitDist * specularWeight;
            sumSpecularWeight += specularWeight;

            float diffuseWeight = w * exponentialWeight(angle, diffuseNormalWeightParam, 0.0f);
            const float sampleDiffuseHitDist = diffuseWeight == 0.0f ? 0.0f : data.y;
            diffuseWeight *= sampleDiffuseHitDist != 0.0f ? 1.0f : 0.0f;
            sumDiffuseHitDist += diffuseWeight == 0.0f ? 0.0f : sampleDiffuseHitDist * diffuseWeight;
            sumDiffuseWeight += diffuseWeight;
        }
    const float4 spec = p.spec.load(px, py), diff = p.diff.load(px, py);
    p.outSpec.store(px, py, make_float4(spec.x, spec.y, spec.z, sumSpecularHitDist / fmaxf(sumSpecularWeight, 1e-6f)));
    p.outDiff.store(px, py, make_float4(diff.x, diff.y, diff.z, sumDiffuseHitDist / fmaxf(sumDiffuseWeight, 1e-6f)));
}

// RELAX_SplitScreen.cs.hlsl:21-62: the noisy input (range-masked; radiance converted to YCoCg in SH mode) left of CommonSettings::splitScreen
struct RelaxSplitScreenParams { TexR32F viewZ; TexRGBA16F diff, spec, diffSh, specSh, outDiff, outSpec, outDiffSh, outSpecSh; };
template <bool SH>
__global__ void __launch_bounds__(BLOCK_W* BLOCK_H) relaxSplitScreenKernel(const __grid_constant__ RelaxConstants cb, const __grid_constant__ RelaxSplitScreenParams p, int ctaY0) {
    pdlEntry();
    const int px = blockIdx.x * BLOCK_W + threadIdx.x, py = (blockIdx.y + ctaY0) * BLOCK_H + threadIdx.y;
    const float u = ((float)px + 0.5f) * cb.rectSizeInv[0];
    if (u > cb.splitScreen || px >= cb.rectSize[0] || py >= cb.rectSize[1]) return;
    const float inRange = relaxInRange(cb, relaxViewZ(cb, p.viewZ.load(px, py))) ? 1.0f : 0.0f;
    const int dx = px >> (cb.diffCheckerboard != 2u ? 1 : 0), sx = px >> (cb.specCheckerboard != 2u ? 1 : 0);
    float4 diff = p.diff.load(dx, py), spec = p.spec.load(sx, py);
    if (SH) {
        diff = f4(linearToYCoCg(xyz(diff)), diff.w);
        spec = f4(linearToYCoCg(xyz(spec)), spec.w);
    }
    p.outDiff.store(px, py, diff * inRange);
    p.outSpec.store(px, py, spec * inRange);
    storeSh<SH>(p.outDiffSh, px, py, loadSh<SH>(p.diffSh, dx, py) * inRange);
    storeSh<SH>(p.outSpecSh, px, py, loadSh<SH>(p.specSh, sx, py) * inRange);
}

struct AtrousTexel {
    float4 spec, diff, nr;
    float3 specSh, diffSh, worldPos;
    float materialID;
};
template <bool SH, class PARAMS>
NRD_DEV AtrousTexel atrousFetch(const RelaxConstants& cb, const PARAMS& p, int x, int y) {
    const int gx = clampi(x, 0, cb.rectSize[0] - 1), gy = clampi(y, 0, cb.rectSize[1] - 1);
    AtrousTexel r;
    r.spec = p.spec.load(gx, gy);
    r.diff = p.diff.load(gx, gy);
    r.specSh = loadSh<SH>(p.specSh, gx, gy);
    r.diffSh = loadSh<SH>(p.diffSh, gx, gy);
    r.nr = unpackNormalRoughness(p.normalRoughness.loadRaw(gx, gy), r.materialID);
    r.worldPos = currentWorldPosPixel(cb, gx, gy, relaxViewZ(cb, p.viewZ.load(gx, gy)));
    return r;
}

constexpr int AT_BORDER = 2, AT_TILE_W = BLOCK_W + 2 * AT_BORDER, AT_TILE_H = BLOCK_H + 2 * AT_BORDER;
#ifndef RELAX_ATROUS_SMEM_MIN_BLOCKS
#define RELAX_ATROUS_SMEM_MIN_BLOCKS 4
#endif
// RES ( all three a-trous kernels ): RelaxSettings::enableRoughnessEdgeStopping, a compile-time switch — the taps evaluate only the specular weight the setting
// selects ( with the constant-buffer flag tested per tap both forms were computed and one discarded: ~4 % of the instructions of these issue-bound launches )
template <bool SH, int SIGNAL, bool RES>
__global__ void __launch_bounds__(BLOCK_W* BLOCK_H, RELAX_ATROUS_SMEM_MIN_BLOCKS) relaxAtrousSmemKernel(const __grid_constant__ RelaxConstants cb, const __grid_constant__ RelaxAtrousParamsT<SIGNAL> p, int ctaY0) {
    pdlEntry();
    // the reference's groupshared tile: 36x12 texels of { illumination + 2nd moment, SH1, normal + roughness, world position + material }
    __shared__ float4 sSpec[AT_TILE_H][AT_TILE_W], sDiff[AT_TILE_H][AT_TILE_W], sNr[AT_TILE_H][AT_TILE_W], sPosMat[AT_TILE_H][AT_TILE_W];
    __shared__ float4 sSpecSh[SH ? AT_TILE_H : 1][SH ? AT_TILE_W : 1], sDiffSh[SH ? AT_TILE_H : 1][SH ? AT_TILE_W : 1];
    const int px = blockIdx.x * BLOCK_W + threadIdx.x, py = (blockIdx.y + ctaY0) * BLOCK_H + threadIdx.y;
    const float skyL = p.tiles.load((blockIdx.x * BLOCK_W) >> 4, py >> 4), skyR = p.tiles.load((blockIdx.x * BLOCK_W + 16) >> 4, py >> 4);
    const float isSky = threadIdx.x < 16 ? skyL : skyR;
    const float viewZpacked = p.viewZ.load(px, py);
    // the prev-frame planes are written by EVERY thread of the reference's 8x8 groups, i.e. up to the rect rounded up to 8 ( this CTA is 32 wide: with
    // dynamic resolution the texels beyond that belong to nobody and stay untouched, as in the reference )
    const bool inReferenceGrid = px < ((cb.rectSize[0] + 7) & ~7);
    if (inReferenceGrid) p.outViewZ.store(px, py, viewZpacked);
    const bool skipTile = skyL != 0.0f && skyR != 0.0f;  // nothing to filter in this CTA: only the prev-frame planes are written
    if (!skipTile) {
        const int baseX = blockIdx.x * BLOCK_W - AT_BORDER, baseY = (blockIdx.y + ctaY0) * BLOCK_H - AT_BORDER;
        for (int i = threadIdx.y * BLOCK_W + threadIdx.x; i < AT_TILE_W * AT_TILE_H; i += BLOCK_W * BLOCK_H) {
            const int tx = i % AT_TILE_W, ty = i / AT_TILE_W;
            const AtrousTexel t = atrousFetch<SH>(cb, p, baseX + tx, baseY + ty);
            sSpec[ty][tx] = t.spec;
            sDiff[ty][tx] = t.diff;
            sNr[ty][tx] = t.nr;
            sPosMat[ty][tx] = f4(t.worldPos, t.materialID);
            if constexpr (SH) {
                sSpecSh[ty][tx] = f4(t.specSh, 0.0f);
                sDiffSh[ty][tx] = f4(t.diffSh, 0.0f);
            }
        }
    }
    __syncthreads();
    const int smx = threadIdx.x + AT_BORDER, smy = threadIdx.y + AT_BORDER;
    auto texel = [&](int dx, int dy) -> AtrousTexel {
        AtrousTexel t;
        t.spec = sSpec[smy + dy][smx + dx];
        t.diff = sDiff[smy + dy][smx + dx];
        t.nr = sNr[smy + dy][smx + dx];
        const float4 pm = sPosMat[smy + dy][smx + dx];
        t.worldPos = xyz(pm);
        t.materialID = pm.w;
        if constexpr (SH) {
            t.specSh = xyz(sSpecSh[smy + dy][smx + dx]);
            t.diffSh = xyz(sDiffSh[smy + dy][smx + dx]);
        } else {
            t.specSh = t.diffSh = f3(0.0f);
        }
        return t;
    };
    const AtrousTexel ctr = skipTile ? atrousFetch<SH>(cb, p, px, py) : texel(0, 0);
    float4 normalRoughness = ctr.nr;
    const float centerViewZ = relaxViewZ(cb, viewZpacked);
    if (!relaxInRange(cb, centerViewZ)) normalRoughness = f4(1.0f / 255.0f);
    if (inReferenceGrid) p.outNormalRoughness.store(px, py, packPrevNormalRoughness(normalRoughness));
    const float3 centerWorldPos = ctr.worldPos;
    const float centerMaterialID = ctr.materialID;
    if (inReferenceGrid) p.outMaterialID.store(px, py, centerMaterialID / 255.0f);

    if (isSky != 0.0f || px >= cb.rectSize[0] || py >= cb.rectSize[1]) return;
    if (!relaxInRange(cb, centerViewZ)) return;

    const float3 centerNormal = xyz(normalRoughness);
    const float centerRoughness = normalRoughness.w;
    const float historyLength = 255.0f * p.historyLength.load(px, py);
    const float kGauss[2] = {0.44198f, 0.27901f};

    if (historyLength >= cb.historyThreshold) {
        const float kernel[2][2] = {{1.0f / 4.0f, 1.0f / 8.0f}, {1.0f / 8.0f, 1.0f / 16.0f}};
        float4 specularSumV = f4(0.0f), diffuseSumV = f4(0.0f);
        // two sweeps over the 3x3 neighbourhood: variance first, then the filter
#pragma unroll
        for (int dx = -1; dx <= 1; dx++)
#pragma unroll
            for (int dy = -1; dy <= 1; dy++) {
                const float k = kernel[dx < 0 ? -dx : dx][dy < 0 ? -dy : dy];
                specularSumV += sSpec[smy + dy][smx + dx] * k;
                diffuseSumV += sDiff[smy + dy][smx + dx] * k;
            }
        const float s1 = luminance(xyz(specularSumV)), d1 = luminance(xyz(diffuseSumV));
        const float centerSpecularVar = fmaxf(0.0f, specularSumV.w - s1 * s1), centerDiffuseVar = fmaxf(0.0f, diffuseSumV.w - d1 * d1);

        const float centerSpecularLuminance = luminance(xyz(ctr.spec));
        const float specularPhiLIlluminationInv = 1.0f / fmaxf(1.0e-4f, cb.specPhiLuminance * sqrtf(centerSpecularVar));
        const float2 roughnessWeightP = roughnessWeightParams(centerRoughness, cb.roughnessFraction);
        const float specularReprojectionConfidence = p.specReprojectionConfidence.load(px, py);
        float specularLuminanceWeightRelaxation = lerp(1.0f, specularReprojectionConfidence, cb.luminanceEdgeStoppingRelaxation);
        float diffuseLobeAngleFraction = cb.lobeAngleFraction, diffuseLobeAngleFractionForSimplifiedSpecularNormalWeight = cb.lobeAngleFraction;
        float specularLobeAngleFraction = cb.lobeAngleFraction, diffuseLuminanceWeightRelaxation = 1.0f;
        if (cb.hasHistoryConfidence) {
            const float2 pixelUv = make_float2((float)px + 0.5f, (float)py + 0.5f) * make_float2(cb.rectSizeInv[0], cb.rectSizeInv[1]);
            const float2 rs = confidenceDrivenRelaxation(cb, p.specConfDummy, pixelUv), rd = confidenceDrivenRelaxation(cb, p.diffConfDummy, pixelUv);
            diffuseLobeAngleFractionForSimplifiedSpecularNormalWeight = lerp(diffuseLobeAngleFraction, 1.0f, rs.x);
            specularLobeAngleFraction = lerp(specularLobeAngleFraction, 1.0f, rs.x);
            specularLuminanceWeightRelaxation *= 1.0f - rs.y;
            diffuseLobeAngleFraction = lerp(diffuseLobeAngleFraction, 1.0f, rd.x);
            diffuseLuminanceWeightRelaxation = 1.0f - rd.y;
        }
        const float specularNormalWeightParamSimplified = normalWeightParam2(1.0f, diffuseLobeAngleFractionForSimplifiedSpecularNormalWeight);
        const float2 specularNormalWeightP = normalWeightParamsAtrous(centerRoughness, historyLength, specularReprojectionConfidence, cb.normalEdgeStoppingRelaxation, specularLobeAngleFraction,
                                                                      cb.specLobeAngleSlack);
        const float3 centerV = -normalize(centerWorldPos);
        const float centerDiffuseLuminance = luminance(xyz(ctr.diff));
        const float diffusePhiLIlluminationInv = 1.0f / fmaxf(1.0e-4f, cb.diffPhiLuminance * sqrtf(centerDiffuseVar));
        const float diffuseNormalWeightParam = normalWeightParam2(1.0f, diffuseLobeAngleFraction);
        const float depthThreshold = cb.depthThreshold * (cb.orthoMode == 0.0f ? centerViewZ : 1.0f);

        float sumWSpecular = 0.0f, sumWDiffuse = 0.0f;
        float4 sumSpecular = f4(0.0f), sumDiffuse = f4(0.0f);
        float3 sumSpecularSH = f3(0.0f), sumDiffuseSH = f3(0.0f);
#pragma unroll
        for (int j = -1; j <= 1; j++)
#pragma unroll
            for (int i = -1; i <= 1; i++) {
                const int x = px + i, y = py + j;
                const bool isCenter = i == 0 && j == 0;
                const bool isInside = x >= 0 && y >= 0 && x < cb.rectSize[0] && y < cb.rectSize[1];
                const float kernelW = isInside ? kGauss[i < 0 ? -i : i] * kGauss[j < 0 ? -j : j] : 0.0f;
                const AtrousTexel s = texel(i, j);
                const float3 sampleNormal = xyz(s.nr);
                float geometryW = planeDistanceWeightAtrous(centerWorldPos, centerNormal, s.worldPos, depthThreshold);
                geometryW *= kernelW;

                const float angles = acosApproxPositive(dot(centerNormal, sampleNormal));
                const float3 sampleV = -normalize(s.worldPos + cb.roughnessEdgeStoppingRelaxation * centerWorldPos);
                const float normalWSpecularSimplified = computeWeight(angles, specularNormalWeightParamSimplified, 0.0f);
                const float normalWSpecular = specularNormalWeightAtrous(specularNormalWeightP, centerNormal, sampleNormal, centerV, sampleV);
                const float roughnessWSpecular = computeWeight(s.nr.w, roughnessWeightP.x, roughnessWeightP.y);
                float specularLuminanceW = fabsf(centerSpecularLuminance - luminance(xyz(s.spec))) * specularPhiLIlluminationInv;
                specularLuminanceW = fminf(cb.specMaxLuminanceRelativeDifference, specularLuminanceW);
                specularLuminanceW *= specularLuminanceWeightRelaxation;
                float wSpecular = geometryW * expf(-specularLuminanceW);
                wSpecular *= RES ? (normalWSpecular * roughnessWSpecular) : normalWSpecularSimplified;
                wSpecular *= compareMaterials(s.materialID, centerMaterialID, cb.specMinMaterial) ? 1.0f : 0.0f;
                wSpecular = isCenter ? kernelW : wSpecular;
                sumWSpecular += wSpecular;
                sumSpecular += wSpecular * s.spec;
                sumSpecularSH += wSpecular * s.specSh;

                const float normalWDiffuse = computeWeight(angles, diffuseNormalWeightParam, 0.0f);
                float diffuseLuminanceW = fabsf(centerDiffuseLuminance - luminance(xyz(s.diff))) * diffusePhiLIlluminationInv;
                diffuseLuminanceW = fminf(cb.diffMaxLuminanceRelativeDifference, diffuseLuminanceW);
                diffuseLuminanceW *= diffuseLuminanceWeightRelaxation;
                float wDiffuse = geometryW * normalWDiffuse * expf(-diffuseLuminanceW);
                wDiffuse *= compareMaterials(s.materialID, centerMaterialID, cb.diffMinMaterial) ? 1.0f : 0.0f;
                wDiffuse = isCenter ? kernelW : wDiffuse;
                sumWDiffuse += wDiffuse;
                sumDiffuse += wDiffuse * s.diff;
                sumDiffuseSH += wDiffuse * s.diffSh;
            }
        sumWSpecular = fmaxf(sumWSpecular, 1e-6f);
        sumSpecular = sumSpecular / sumWSpecular;
        const float sp1 = luminance(xyz(sumSpecular));
        p.outSpec.store(px, py, f4(xyz(sumSpecular), fmaxf(0.0f, sumSpecular.w - sp1 * sp1)));
        storeSh<SH>(p.outSpecSh, px, py, sumSpecularSH / sumWSpecular);
        sumWDiffuse = fmaxf(sumWDiffuse, 1e-6f);
        sumDiffuse = sumDiffuse / sumWDiffuse;
        const float dp1 = luminance(xyz(sumDiffuse));
        p.outDiff.store(px, py, f4(xyz(sumDiffuse), fmaxf(0.0f, sumDiffuse.w - dp1 * dp1)));
        storeSh<SH>(p.outDiffSh, px, py, sumDiffuseSH / sumWDiffuse);
    } else {
        float sumWS = 0.0f, sumS1 = 0.0f, sumS2 = 0.0f, sumWD = 0.0f, sumD1 = 0.0f, sumD2 = 0.0f;
        float3 sumS = f3(0.0f), sumD = f3(0.0f), sumSSH = f3(0.0f), sumDSH = f3(0.0f);
        const float diffuseNormalWeightParam = normalWeightParam2(1.0f, cb.lobeAngleFraction);
        for (int cx = -2; cx <= 2; cx++)
            for (int cy = -2; cy <= 2; cy++) {
                const AtrousTexel s = texel(cx, cy);
                const float normalW = computeWeight(acosApproxPositive(dot(centerNormal, xyz(s.nr))), diffuseNormalWeightParam, 0.0f);
                const float specularW = normalW * (compareMaterials(s.materialID, centerMaterialID, cb.specMinMaterial) ? 1.0f : 0.0f);
                sumWS += specularW;
                sumS += xyz(s.spec) * specularW;
                sumS1 += luminance(xyz(s.spec)) * specularW;
                sumS2 += s.spec.w * specularW;
                sumSSH += s.specSh * specularW;
                const float diffuseW = normalW * (compareMaterials(s.materialID, centerMaterialID, cb.diffMinMaterial) ? 1.0f : 0.0f);
                sumWD += diffuseW;
                sumD += xyz(s.diff) * diffuseW;
                sumD1 += luminance(xyz(s.diff)) * diffuseW;
                sumD2 += s.diff.w * diffuseW;
                sumDSH += s.diffSh * diffuseW;
            }
        const float boost = fmaxf(1.0f, 4.0f / (historyLength + 1.0f));
        sumWS = fmaxf(sumWS, 1e-6f);
        sumS = sumS / sumWS;
        sumS1 /= sumWS;
        sumS2 /= sumWS;
        p.outSpec.store(px, py, f4(sumS, fmaxf(0.0f, sumS2 - sumS1 * sumS1) * boost));
        storeSh<SH>(p.outSpecSh, px, py, sumSSH / sumWS);
        sumWD = fmaxf(sumWD, 1e-6f);
        sumD = sumD / sumWD;
        sumD1 /= sumWD;
        sumD2 /= sumWD;
        p.outDiff.store(px, py, f4(sumD, fmaxf(0.0f, sumD2 - sumD1 * sumD1) * boost));
        storeSh<SH>(p.outDiffSh, px, py, sumDSH / sumWD);
    }
}

#ifndef RELAX_ATROUS_GATHER_MIN_BLOCKS
#define RELAX_ATROUS_GATHER_MIN_BLOCKS 4
#endif
template <bool SH, int SIGNAL, bool RES>
__global__ void __launch_bounds__(BLOCK_W* BLOCK_H, RELAX_ATROUS_GATHER_MIN_BLOCKS) relaxAtrousKernel(const __grid_constant__ RelaxConstants cb, const __grid_constant__ RelaxAtrousParamsT<SIGNAL> p, int ctaY0) {
    pdlEntry();
    const int px = blockIdx.x * BLOCK_W + threadIdx.x, py = (blockIdx.y + ctaY0) * BLOCK_H + threadIdx.y;
    if (p.tiles.load(px >> 4, py >> 4) != 0.0f || px >= cb.rectSize[0] || py >= cb.rectSize[1]) return;
    const float centerViewZ = relaxViewZ(cb, p.viewZ.load(px, py));
    if (!relaxInRange(cb, centerViewZ)) return;

    float centerMaterialID;
    const float4 centerNormalRoughness = unpackNormalRoughness(p.normalRoughness.loadRaw(px, py), centerMaterialID);
    const float3 centerNormal = xyz(centerNormalRoughness);
    const float centerRoughness = centerNormalRoughness.w;
    const float historyLength = 255.0f * p.historyLength.load(px, py);
    const float stepSize = (float)cb.stepSize;
    const float kGauss[2] = {0.44198f, 0.27901f};

    float diffuseLobeAngleFraction = (SH ? 1.0f : cb.lobeAngleFraction) / sqrtf(stepSize);  // RELAX_Atrous.cs.hlsl:46-49
    diffuseLobeAngleFraction = lerp(0.99f, diffuseLobeAngleFraction, saturate(historyLength / 5.0f));

    const float4 centerSpecular = p.spec.load(px, py);
    const float centerSpecularLuminance = luminance(xyz(centerSpecular));
    const float specularPhiLIlluminationInv = 1.0f / fmaxf(1.0e-4f, cb.specPhiLuminance * sqrtf(centerSpecular.w));
    const float2 roughnessWeightP = roughnessWeightParams(centerRoughness, cb.roughnessFraction);
    const float specularReprojectionConfidence = p.specReprojectionConfidence.load(px, py);
    float specularLuminanceWeightRelaxation = 1.0f;
    if (cb.stepSize <= 4) specularLuminanceWeightRelaxation = lerp(1.0f, specularReprojectionConfidence, cb.luminanceEdgeStoppingRelaxation);
    float diffuseLobeAngleFractionForSimplifiedSpecularNormalWeight = diffuseLobeAngleFraction, specularLobeAngleFraction = cb.lobeAngleFraction;
    float diffuseLuminanceWeightRelaxation = 1.0f;
    if (cb.hasHistoryConfidence) {
        const float2 pixelUv = make_float2((float)px + 0.5f, (float)py + 0.5f) * make_float2(cb.rectSizeInv[0], cb.rectSizeInv[1]);
        const float2 rs = confidenceDrivenRelaxation(cb, p.specConfDummy, pixelUv), rd = confidenceDrivenRelaxation(cb, p.diffConfDummy, pixelUv);
        diffuseLobeAngleFractionForSimplifiedSpecularNormalWeight = lerp(diffuseLobeAngleFraction, 1.0f, rs.x);
        specularLobeAngleFraction = lerp(specularLobeAngleFraction, 1.0f, rs.x);
        specularLuminanceWeightRelaxation *= 1.0f - rs.y;
        diffuseLobeAngleFraction = lerp(diffuseLobeAngleFraction, 1.0f, rd.x);
        diffuseLuminanceWeightRelaxation = 1.0f - rd.y;
    }
    const float specularNormalWeightParamSimplified = normalWeightParam2(1.0f, diffuseLobeAngleFractionForSimplifiedSpecularNormalWeight);
    const float2 specularNormalWeightP = normalWeightParamsAtrous(centerRoughness, historyLength, specularReprojectionConfidence, cb.normalEdgeStoppingRelaxation, specularLobeAngleFraction,
                                                                  cb.specLobeAngleSlack);
    float sumWSpecular = 0.44198f * 0.44198f;
    float4 sumSpecular = centerSpecular * make_float4(sumWSpecular, sumWSpecular, sumWSpecular, sumWSpecular * sumWSpecular);
    float3 sumSpecularSH = loadSh<SH>(p.specSh, px, py) * sumWSpecular;

    const float4 centerDiffuse = p.diff.load(px, py);
    const float centerDiffuseLuminance = luminance(xyz(centerDiffuse));
    const float diffusePhiLIlluminationInv = 1.0f / fmaxf(1.0e-4f, cb.diffPhiLuminance * sqrtf(centerDiffuse.w));
    const float diffuseNormalWeightParam = normalWeightParam2(1.0f, diffuseLobeAngleFraction);
    float sumWDiffuse = 0.44198f * 0.44198f;
    float4 sumDiffuse = centerDiffuse * make_float4(sumWDiffuse, sumWDiffuse, sumWDiffuse, sumWDiffuse * sumWDiffuse);
    float3 sumDiffuseSH = loadSh<SH>(p.diffSh, px, py) * sumWDiffuse;

    const float3 centerWorldPos = currentWorldPosPixel(cb, px, py, centerViewZ);
    const float3 centerV = -normalize(centerWorldPos);
    const float depthThreshold = cb.depthThreshold * (cb.orthoMode == 0.0f ? centerViewZ : 1.0f);

    int offX = 0, offY = 0;
    if (cb.stepSize > 4) {
        Rng rng;
        rng.init((uint32_t)px, (uint32_t)py, cb.frameIndex);
        const float rx = rng.next(), ry = rng.next();
        offX = (int)(stepSize * 0.5f * (rx - 0.5f));
        offY = (int)(stepSize * 0.5f * (ry - 0.5f));
    }
#pragma unroll
    for (int j = -1; j <= 1; j++)
#pragma unroll
        for (int i = -1; i <= 1; i++) {
            if (i == 0 && j == 0) continue;
            const int x = px + offX + i * (int)cb.stepSize, y = py + offY + j * (int)cb.stepSize;
            const bool isInside = x >= 0 && y >= 0 && x < cb.rectSize[0] && y < cb.rectSize[1];
            const float kernelW = kGauss[i < 0 ? -i : i] * kGauss[j < 0 ? -j : j];

            float sampleMaterialID;
            const float4 sampleNormalRoughness = unpackNormalRoughness(p.normalRoughness.loadRaw(x, y), sampleMaterialID);
            const float3 sampleNormal = xyz(sampleNormalRoughness);
            const float sampleViewZ = relaxViewZ(cb, p.viewZ.load(x, y));
            const float3 sampleWorldPos = currentWorldPosPixel(cb, x, y, sampleViewZ);
            float geometryW = planeDistanceWeightAtrous(centerWorldPos, centerNormal, sampleWorldPos, depthThreshold);
            geometryW *= kernelW;
            geometryW *= (isInside && relaxInRange(cb, sampleViewZ)) ? 1.0f : 0.0f;

            const float3 sampleV = -normalize(sampleWorldPos + cb.roughnessEdgeStoppingRelaxation * centerWorldPos);
            const float angles = acosApproxPositive(dot(centerNormal, sampleNormal));
            const float normalWSpecularSimplified = computeWeight(angles, specularNormalWeightParamSimplified, 0.0f);
            const float normalWSpecular = specularNormalWeightAtrous(specularNormalWeightP, centerNormal, sampleNormal, centerV, sampleV);
            const float roughnessWSpecular = computeWeight(sampleNormalRoughness.w, roughnessWeightP.x, roughnessWeightP.y);
            float wSpecular = geometryW * (RES ? (normalWSpecular * roughnessWSpecular) : normalWSpecularSimplified);
            wSpecular *= compareMaterials(sampleMaterialID, centerMaterialID, cb.specMinMaterial) ? 1.0f : 0.0f;
            // The shader reads a tap's radiance / SH only `if( w > 1e-4 )`. Here the four texels are requested together with the tap's geometry and a skipped tap
            // contributes exact zeros instead: nothing waits for a weight before its loads go out, so the loads of all eight taps overlap. The gathering
            // launches ( strides 8, 16 ) wait on memory latency, not on issue slots: 268 / 274 -> 236 / 241 us per launch, sums bit-identical.
            const int lx = clampi(x, 0, cb.rectSize[0] - 1), ly = clampi(y, 0, cb.rectSize[1] - 1);
            {
                const bool use = wSpecular > 1e-4f;
                float4 s = p.spec.fetch(lx, ly);
                float3 ssh = fetchSh<SH>(p.specSh, lx, ly);
                if (!use) { s = f4(0.0f); ssh = f3(0.0f); }
                float lw = fabsf(centerSpecularLuminance - luminance(xyz(s))) * specularPhiLIlluminationInv;
                lw = fminf(cb.specMaxLuminanceRelativeDifference, lw);
                lw *= specularLuminanceWeightRelaxation;
                wSpecular = use ? wSpecular * expf(-lw) : 0.0f;
                sumWSpecular += wSpecular;
                sumSpecular += make_float4(wSpecular, wSpecular, wSpecular, wSpecular * wSpecular) * s;
                sumSpecularSH += ssh * wSpecular;
            }

            const float normalWDiffuse = computeWeight(angles, diffuseNormalWeightParam, 0.0f);
            float wDiffuse = geometryW * normalWDiffuse;
            wDiffuse *= compareMaterials(sampleMaterialID, centerMaterialID, cb.diffMinMaterial) ? 1.0f : 0.0f;
            {
                const bool use = wDiffuse > 1e-4f;
                float4 s = p.diff.fetch(lx, ly);
                float3 ssh = fetchSh<SH>(p.diffSh, lx, ly);
                if (!use) { s = f4(0.0f); ssh = f3(0.0f); }
                float lw = fabsf(centerDiffuseLuminance - luminance(xyz(s))) * diffusePhiLIlluminationInv;
                lw = fminf(cb.diffMaxLuminanceRelativeDifference, lw);
                lw *= diffuseLuminanceWeightRelaxation;
                wDiffuse = use ? wDiffuse * expf(-lw) : 0.0f;
                sumWDiffuse += wDiffuse;
                sumDiffuse += make_float4(wDiffuse, wDiffuse, wDiffuse, wDiffuse * wDiffuse) * s;
                sumDiffuseSH += ssh * wDiffuse;
            }
        }
    const float currHistoryLength = fmaxf(historyLength - 1.0f, 0.0f);
    float4 filteredSpecular = sumSpecular / make_float4(sumWSpecular, sumWSpecular, sumWSpecular, sumWSpecular * sumWSpecular);
    if (cb.isLastPass == 1) filteredSpecular = f4(SH ? linearToYCoCg(xyz(filteredSpecular)) : xyz(filteredSpecular), currHistoryLength);   // YCoCg output in SH mode only
    storeSh<SH>(p.outSpecSh, px, py, sumSpecularSH / sumWSpecular);
    p.outSpec.store(px, py, filteredSpecular);
    float4 filteredDiffuse = sumDiffuse / make_float4(sumWDiffuse, sumWDiffuse, sumWDiffuse, sumWDiffuse * sumWDiffuse);
    if (cb.isLastPass == 1) filteredDiffuse = f4(SH ? linearToYCoCg(xyz(filteredDiffuse)) : xyz(filteredDiffuse), currHistoryLength);
    storeSh<SH>(p.outDiffSh, px, py, sumDiffuseSH / sumWDiffuse);
    p.outDiff.store(px, py, filteredDiffuse);
}

// ---------------------------------------------------------------------------------------------------------------
// TMA ( cp.async.bulk.tensor ) for the tiles that are staged RAW. The radiance / SH neighbourhoods of the tiled a-trous kernel are the textures' own fp16 words
// ( they are converted per tap, where the gathering kernel converts what it loaded ), so one elected thread asks the TMA unit for the ( 32 + 2 STEP ) x ( 8 + 2 STEP )
// box of each of the four RGBA16F textures and the other 255 threads spend no LDG / address / STS instructions on them; the boxes land while the CTA decodes the
// { normal, roughness } / { world position, material } / viewZ arrays, which cannot come from TMA ( they are computed, not copied ). Texels of a box outside the
// texture are zero-filled instead of clamped: a tap outside the rect has weight 0 and never reads them. A tensor map describes one texture as 2 x width uint32
// elements per row ( an RGBA16F texel = two of them ); it is a kernel PARAMETER like the views, so cached CUDA graphs patch it per frame like everything else.
struct alignas(64) RelaxAtrousTma { CUtensorMap spec, diff, specSh, diffSh; };
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn tensorMapEncoder() {
    static EncodeTiledFn fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
            (void)cudaGetLastError();
            f = nullptr;
        }
        return (EncodeTiledFn)f;
    }();
    return fn;
}
// false: the texture cannot be described ( base not 16-byte aligned, pitch not a multiple of 16 bytes ) or the driver has no encoder — the caller keeps the LDG staging
inline bool encodeRgba16fBox(CUtensorMap& map, const TexView& t, int boxW, int boxH) {
    EncodeTiledFn encode = tensorMapEncoder();
    if (!encode || !t.data || ((uintptr_t)t.data & 15u) != 0 || ((size_t)t.pitch * 8u) % 16u != 0 || t.w <= 0 || t.h <= 0) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)t.w * 2u, (cuuint64_t)t.h};
    const cuuint64_t strides[1] = {(cuuint64_t)t.pitch * 8u};
    const cuuint32_t box[2] = {(cuuint32_t)boxW * 2u, (cuuint32_t)boxH}, elem[2] = {1u, 1u};
    return encode(&map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, t.data, dims, strides, box, elem, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
NRD_DEV uint32_t smemAddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
NRD_DEV void mbarrierInit(uint64_t* bar, uint32_t arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(arrivals) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");   // the async proxy ( TMA ) must see the initialised barrier
}
NRD_DEV void mbarrierExpectTx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes) : "memory"); }
NRD_DEV void mbarrierWait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred done;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 done, [%0], %1;\n"
        "@done bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smemAddr(bar)), "r"(parity)
        : "memory");
}
// box of `map` whose first element is ( x, y ) — in uint32 elements / rows, may be negative or reach past the texture — into shared memory at dst ( 128-byte aligned )
NRD_DEV void tmaLoadBox2D(void* dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smemAddr(dst)), "l"(map), "r"(smemAddr(bar)), "r"(x), "r"(y)
                 : "memory");
}

// The same pass for the small strides ( STEP = 2, 4 ) with the neighbourhood staged in shared memory: a 32x8 CTA needs ( 32 + 2 STEP ) x ( 8 + 2 STEP ) texels, 1.7x / 2.5x
// its own pixels, and every one of them is fetched, unpacked ( normal + roughness + material ) and turned into a world position ONCE instead of once per tap that
// lands on it ( 8 taps per pixel ): ~80 of the ~360 instructions of a tap are that decode. The radiance / SH texels are staged as the raw fp16 words ( converted per
// tap exactly where the gathering kernel converts what it loaded ), so the tile is 68 B per texel: 29 KB ( STEP 2 ) / 43 KB ( STEP 4 ). The arithmetic per tap is the
// gathering kernel's, character for character: results are bit-identical. Strides 8 and 16 ( 4.5x / 10x the pixels, random tap offsets ) keep gathering.
// TMA: the four raw planes arrive through cp.async.bulk.tensor ( above ) instead of LDG + STS
template <bool SH, int SIGNAL, int STEP, bool TMA, bool RES>
__global__ void __launch_bounds__(BLOCK_W* BLOCK_H, RELAX_ATROUS_SMEM_MIN_BLOCKS) relaxAtrousTiledKernel(const __grid_constant__ RelaxConstants cb, const __grid_constant__ RelaxAtrousParamsT<SIGNAL> p, int ctaY0,
                                                                                                      const __grid_constant__ RelaxAtrousTma tma) {
    pdlEntry();
    constexpr bool HAS_SPEC = (SIGNAL & SIGNAL_SPEC) != 0, HAS_DIFF = (SIGNAL & SIGNAL_DIFF) != 0;
    constexpr int TW = BLOCK_W + 2 * STEP, TH = BLOCK_H + 2 * STEP;
    static_assert((TW * TH * 8) % 128 == 0, "a TMA box lands on a 128-byte aligned address");
    __shared__ float4 sNr[TH][TW], sPosMat[TH][TW];   // { normal, roughness }, { world position, material }
    __shared__ float sZ[TH][TW];
    __shared__ alignas(128) uint2 sSpec[HAS_SPEC ? TH : 1][HAS_SPEC ? TW : 1];
    __shared__ alignas(128) uint2 sDiff[HAS_DIFF ? TH : 1][HAS_DIFF ? TW : 1];
    __shared__ alignas(128) uint2 sSpecSh[(HAS_SPEC && SH) ? TH : 1][(HAS_SPEC && SH) ? TW : 1];
    __shared__ alignas(128) uint2 sDiffSh[(HAS_DIFF && SH) ? TH : 1][(HAS_DIFF && SH) ? TW : 1];
    __shared__ uint64_t sTmaBar;
    const int px = blockIdx.x * BLOCK_W + threadIdx.x, py = (blockIdx.y + ctaY0) * BLOCK_H + threadIdx.y;
    // the CTA covers two 16x16 tiles of one tile row
    const float skyL = p.tiles.load((blockIdx.x * BLOCK_W) >> 4, py >> 4), skyR = p.tiles.load((blockIdx.x * BLOCK_W + 16) >> 4, py >> 4);
    if (skyL != 0.0f && skyR != 0.0f) return;
    {
        const int baseX = blockIdx.x * BLOCK_W - STEP, baseY = (blockIdx.y + ctaY0) * BLOCK_H - STEP;
        if constexpr (TMA) {
            if (threadIdx.x == 0 && threadIdx.y == 0) {
                constexpr uint32_t kPlanes = (HAS_SPEC ? 1u : 0u) + (HAS_DIFF ? 1u : 0u) + ((HAS_SPEC && SH) ? 1u : 0u) + ((HAS_DIFF && SH) ? 1u : 0u);
                mbarrierInit(&sTmaBar, 1u);
                mbarrierExpectTx(&sTmaBar, kPlanes * (uint32_t)(TW * TH * 8));
                if constexpr (HAS_SPEC) tmaLoadBox2D(&sSpec[0][0], &tma.spec, baseX * 2, baseY, &sTmaBar);
                if constexpr (HAS_DIFF) tmaLoadBox2D(&sDiff[0][0], &tma.diff, baseX * 2, baseY, &sTmaBar);
                if constexpr (HAS_SPEC && SH) tmaLoadBox2D(&sSpecSh[0][0], &tma.specSh, baseX * 2, baseY, &sTmaBar);
                if constexpr (HAS_DIFF && SH) tmaLoadBox2D(&sDiffSh[0][0], &tma.diffSh, baseX * 2, baseY, &sTmaBar);
            }
        }
        for (int i = threadIdx.y * BLOCK_W + threadIdx.x; i < TW * TH; i += BLOCK_W * BLOCK_H) {
            const int tx = i % TW, ty = i / TW;
            // clamped to the rect: a tap outside it has weight 0 and never reads what is staged here for it
            const int gx = clampi(baseX + tx, 0, cb.rectSize[0] - 1), gy = clampi(baseY + ty, 0, cb.rectSize[1] - 1);
            float materialID;
            const float4 nr = unpackNormalRoughness(p.normalRoughness.fetchRaw(gx, gy), materialID);
            const float z = relaxViewZ(cb, p.viewZ.fetch(gx, gy));
            sNr[ty][tx] = nr;
            sZ[ty][tx] = z;
            sPosMat[ty][tx] = f4(currentWorldPosPixel(cb, gx, gy, z), materialID);
            if constexpr (!TMA) {
                if constexpr (HAS_SPEC) sSpec[ty][tx] = p.spec.fetchRaw(gx, gy);
                if constexpr (HAS_DIFF) sDiff[ty][tx] = p.diff.fetchRaw(gx, gy);
                if constexpr (HAS_SPEC && SH) sSpecSh[ty][tx] = p.specSh.fetchRaw(gx, gy);
                if constexpr (HAS_DIFF && SH) sDiffSh[ty][tx] = p.diffSh.fetchRaw(gx, gy);
            }
        }
    }
    __syncthreads();
    // every thread waits for the boxes, also the ones about to return: the CTA's shared memory must not be handed to the next CTA with a copy in flight
    if constexpr (TMA) mbarrierWait(&sTmaBar, 0u);
    if ((threadIdx.x < 16 ? skyL : skyR) != 0.0f || px >= cb.rectSize[0] || py >= cb.rectSize[1]) return;
    const int smx = threadIdx.x + STEP, smy = threadIdx.y + STEP;
    const float centerViewZ = sZ[smy][smx];
    if (!relaxInRange(cb, centerViewZ)) return;
    // the staged texel of a tap / of the centre, decoded like the gathering kernel decodes its loads
    auto tapSpec = [&](int y, int x) -> float4 { if constexpr (HAS_SPEC) return TexRGBA16F::decode(sSpec[y][x]); else return f4(0.0f); };
    auto tapDiff = [&](int y, int x) -> float4 { if constexpr (HAS_DIFF) return TexRGBA16F::decode(sDiff[y][x]); else return f4(0.0f); };
    auto tapSpecSh = [&](int y, int x) -> float3 { if constexpr (HAS_SPEC && SH) return xyz(TexRGBA16F::decode(sSpecSh[y][x])); else return f3(0.0f); };
    auto tapDiffSh = [&](int y, int x) -> float3 { if constexpr (HAS_DIFF && SH) return xyz(TexRGBA16F::decode(sDiffSh[y][x])); else return f3(0.0f); };

    const float4 centerPosMat = sPosMat[smy][smx];
    const float centerMaterialID = centerPosMat.w;
    const float4 centerNormalRoughness = sNr[smy][smx];
    const float3 centerNormal = xyz(centerNormalRoughness);
    const float centerRoughness = centerNormalRoughness.w;
    const float historyLength = 255.0f * p.historyLength.load(px, py);
    const float stepSize = (float)cb.stepSize;
    const float kGauss[2] = {0.44198f, 0.27901f};

    float diffuseLobeAngleFraction = (SH ? 1.0f : cb.lobeAngleFraction) / sqrtf(stepSize);  // RELAX_Atrous.cs.hlsl:46-49
    diffuseLobeAngleFraction = lerp(0.99f, diffuseLobeAngleFraction, saturate(historyLength / 5.0f));

    const float4 centerSpecular = tapSpec(smy, smx);
    const float centerSpecularLuminance = luminance(xyz(centerSpecular));
    const float specularPhiLIlluminationInv = 1.0f / fmaxf(1.0e-4f, cb.specPhiLuminance * sqrtf(centerSpecular.w));
    const float2 roughnessWeightP = roughnessWeightParams(centerRoughness, cb.roughnessFraction);
    const float specularReprojectionConfidence = p.specReprojectionConfidence.load(px, py);
    float specularLuminanceWeightRelaxation = 1.0f;
    if (cb.stepSize <= 4) specularLuminanceWeightRelaxation = lerp(1.0f, specularReprojectionConfidence, cb.luminanceEdgeStoppingRelaxation);
    float diffuseLobeAngleFractionForSimplifiedSpecularNormalWeight = diffuseLobeAngleFraction, specularLobeAngleFraction = cb.lobeAngleFraction;
    float diffuseLuminanceWeightRelaxation = 1.0f;
    if (cb.hasHistoryConfidence) {
        const float2 pixelUv = make_float2((float)px + 0.5f, (float)py + 0.5f) * make_float2(cb.rectSizeInv[0], cb.rectSizeInv[1]);
        const float2 rs = confidenceDrivenRelaxation(cb, p.specConfDummy, pixelUv), rd = confidenceDrivenRelaxation(cb, p.diffConfDummy, pixelUv);
        diffuseLobeAngleFractionForSimplifiedSpecularNormalWeight = lerp(diffuseLobeAngleFraction, 1.0f, rs.x);
        specularLobeAngleFraction = lerp(specularLobeAngleFraction, 1.0f, rs.x);
        specularLuminanceWeightRelaxation *= 1.0f - rs.y;
        diffuseLobeAngleFraction = lerp(diffuseLobeAngleFraction, 1.0f, rd.x);
        diffuseLuminanceWeightRelaxation = 1.0f - rd.y;
    }
    const float specularNormalWeightParamSimplified = normalWeightParam2(1.0f, diffuseLobeAngleFractionForSimplifiedSpecularNormalWeight);
    const float2 specularNormalWeightP = normalWeightParamsAtrous(centerRoughness, historyLength, specularReprojectionConfidence, cb.normalEdgeStoppingRelaxation, specularLobeAngleFraction,
                                                                  cb.specLobeAngleSlack);
    float sumWSpecular = 0.44198f * 0.44198f;
    float4 sumSpecular = centerSpecular * make_float4(sumWSpecular, sumWSpecular, sumWSpecular, sumWSpecular * sumWSpecular);
    float3 sumSpecularSH = tapSpecSh(smy, smx) * sumWSpecular;

    const float4 centerDiffuse = tapDiff(smy, smx);
    const float centerDiffuseLuminance = luminance(xyz(centerDiffuse));
    const float diffusePhiLIlluminationInv = 1.0f / fmaxf(1.0e-4f, cb.diffPhiLuminance * sqrtf(centerDiffuse.w));
    const float diffuseNormalWeightParam = normalWeightParam2(1.0f, diffuseLobeAngleFraction);
    float sumWDiffuse = 0.44198f * 0.44198f;
    float4 sumDiffuse = centerDiffuse * make_float4(sumWDiffuse, sumWDiffuse, sumWDiffuse, sumWDiffuse * sumWDiffuse);
    float3 sumDiffuseSH = tapDiffSh(smy, smx) * sumWDiffuse;

    const float3 centerWorldPos = xyz(centerPosMat);
    const float3 centerV = -normalize(centerWorldPos);
    const float depthThreshold = cb.depthThreshold * (cb.orthoMode == 0.0f ? centerViewZ : 1.0f);

#pragma unroll
    for (int j = -1; j <= 1; j++)
#pragma unroll
        for (int i = -1; i <= 1; i++) {
            if (i == 0 && j == 0) continue;
            const int x = px + i * STEP, y = py + j * STEP;
            const int tx = smx + i * STEP, ty = smy + j * STEP;
            const bool isInside = x >= 0 && y >= 0 && x < cb.rectSize[0] && y < cb.rectSize[1];
            const float kernelW = kGauss[i < 0 ? -i : i] * kGauss[j < 0 ? -j : j];

            const float4 samplePosMat = sPosMat[ty][tx];
            const float sampleMaterialID = samplePosMat.w;
            const float4 sampleNormalRoughness = sNr[ty][tx];
            const float3 sampleNormal = xyz(sampleNormalRoughness);
            const float sampleViewZ = sZ[ty][tx];
            const float3 sampleWorldPos = xyz(samplePosMat);
            float geometryW = planeDistanceWeightAtrous(centerWorldPos, centerNormal, sampleWorldPos, depthThreshold);
            geometryW *= kernelW;
            geometryW *= (isInside && relaxInRange(cb, sampleViewZ)) ? 1.0f : 0.0f;

            const float3 sampleV = -normalize(sampleWorldPos + cb.roughnessEdgeStoppingRelaxation * centerWorldPos);
            const float angles = acosApproxPositive(dot(centerNormal, sampleNormal));
            const float normalWSpecularSimplified = computeWeight(angles, specularNormalWeightParamSimplified, 0.0f);
            const float normalWSpecular = specularNormalWeightAtrous(specularNormalWeightP, centerNormal, sampleNormal, centerV, sampleV);
            const float roughnessWSpecular = computeWeight(sampleNormalRoughness.w, roughnessWeightP.x, roughnessWeightP.y);
            float wSpecular = geometryW * (RES ? (normalWSpecular * roughnessWSpecular) : normalWSpecularSimplified);
            wSpecular *= compareMaterials(sampleMaterialID, centerMaterialID, cb.specMinMaterial) ? 1.0f : 0.0f;
            if (wSpecular > 1e-4f) {
                const float4 s = tapSpec(ty, tx);
                float lw = fabsf(centerSpecularLuminance - luminance(xyz(s))) * specularPhiLIlluminationInv;
                lw = fminf(cb.specMaxLuminanceRelativeDifference, lw);
                lw *= specularLuminanceWeightRelaxation;
                wSpecular *= expf(-lw);
                sumWSpecular += wSpecular;
                sumSpecular += make_float4(wSpecular, wSpecular, wSpecular, wSpecular * wSpecular) * s;
                sumSpecularSH += tapSpecSh(ty, tx) * wSpecular;
            }

            const float normalWDiffuse = computeWeight(angles, diffuseNormalWeightParam, 0.0f);
            float wDiffuse = geometryW * normalWDiffuse;
            wDiffuse *= compareMaterials(sampleMaterialID, centerMaterialID, cb.diffMinMaterial) ? 1.0f : 0.0f;
            if (wDiffuse > 1e-4f) {
                const float4 s = tapDiff(ty, tx);
                float lw = fabsf(centerDiffuseLuminance - luminance(xyz(s))) * diffusePhiLIlluminationInv;
                lw = fminf(cb.diffMaxLuminanceRelativeDifference, lw);
                lw *= diffuseLuminanceWeightRelaxation;
                wDiffuse *= expf(-lw);
                sumWDiffuse += wDiffuse;
                sumDiffuse += make_float4(wDiffuse, wDiffuse, wDiffuse, wDiffuse * wDiffuse) * s;
                sumDiffuseSH += tapDiffSh(ty, tx) * wDiffuse;
            }
        }
    const float currHistoryLength = fmaxf(historyLength - 1.0f, 0.0f);
    float4 filteredSpecular = sumSpecular / make_float4(sumWSpecular, sumWSpecular, sumWSpecular, sumWSpecular * sumWSpecular);
    if (cb.isLastPass == 1) filteredSpecular = f4(SH ? linearToYCoCg(xyz(filteredSpecular)) : xyz(filteredSpecular), currHistoryLength);   // YCoCg output in SH mode only
    storeSh<SH>(p.outSpecSh, px, py, sumSpecularSH / sumWSpecular);
    p.outSpec.store(px, py, filteredSpecular);
    float4 filteredDiffuse = sumDiffuse / make_float4(sumWDiffuse, sumWDiffuse, sumWDiffuse, sumWDiffuse * sumWDiffuse);
    if (cb.isLastPass == 1) filteredDiffuse = f4(SH ? linearToYCoCg(xyz(filteredDiffuse)) : xyz(filteredDiffuse), currHistoryLength);
    storeSh<SH>(p.outDiffSh, px, py, sumDiffuseSH / sumWDiffuse);
    p.outDiff.store(px, py, filteredDiffuse);
}

// ---------------------------------------------------------------------------------------------------------------
// RELAX validation overlay ( RELAX_Validation.cs.hlsl:31-212 ): the viewports of the REBLUR overlay that RELAX has data for — normals, roughness, viewZ,
// motion-vector error, world units + jitter, accumulated frames. See kernels/debug_overlay.cuh for the grid and the captions.
struct RelaxValidationParams {
    TexNR normalRoughness; TexR32F viewZ; TexView mv; TexR8 historyLength; TexView out;
};
__global__ void __launch_bounds__(256) relaxValidationKernel(const __grid_constant__ RelaxConstants cb, const __grid_constant__ RelaxValidationParams p) {
    pdlEntry();
    const int px = blockIdx.x * 32 + (threadIdx.x & 31), py = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (!p.out.inside(px, py)) return;
    if (cb.resetHistory != 0u) {
        anyStore4(p.out, px, py, f4(0.0f));
        return;
    }
    const float2 resourceSize = make_float2(cb.resourceSize[0], cb.resourceSize[1]);
    const OverlayCell cell = overlayCell(px, py, resourceSize);
    const float2 uvScaled = cell.uv * make_float2(cb.resolutionScale[0], cb.resolutionScale[1]);
    const float4 nr = unpackNormalRoughness(p.normalRoughness.sampleNearestRaw(uvScaled));
    const float viewZ = relaxViewZ(cb, p.viewZ.sampleNearest(uvScaled));
    const float4 mvRaw = anyFetch4(p.mv, p.mv.cx((int)floorf(uvScaled.x * (float)p.mv.w)), p.mv.cy((int)floorf(uvScaled.y * (float)p.mv.h)));
    const float3 mv = make_float3(mvRaw.x * cb.mvScale[0], mvRaw.y * cb.mvScale[1], mvRaw.z * cb.mvScale[2]);
    const float historyLength = 255.0f * p.historyLength.sampleNearest(uvScaled) - 1.0f;
    const float3 X = currentWorldPosClip(cb, cell.uv * 2.0f - 1.0f, viewZ);
    const bool isInf = !relaxInRange(cb, viewZ);
    const bool checker = ((((unsigned)px >> 2) ^ ((unsigned)py >> 2)) & 1u) != 0u;
    const float notInf = isInf ? 0.0f : 1.0f;

    Caption text(px, py, cell.captionX, cell.captionY);
    float4 result = anyFetch4(p.out, px, py);
    auto set = [&](float3 c) { result = f4(c, 1.0f); };
    switch (cell.index) {
        case 0:
            text.print("NORMALS");
            text.nextChar();
            text.printUint(2u);
            set(xyz(nr) * 0.5f + 0.5f);
            break;
        case 1:
            text.print("ROUGHNESS");
            text.nextChar();
            text.printUint(1u);
            set(f3(nr.w));
            break;
        case 2: {
            text.print("Z");
            const float f = 0.1f * viewZ / (1.0f + 0.1f * viewZ);
            set(isInf ? make_float3(1.0f, 0.0f, 0.0f) : make_float3(0.0f, f, 0.0f));
            break;
        }
        case 3: {
            text.print("MV");
            const float2 expected = screenUv(cb.worldToClipPrev, X);
            float2 prev = cell.uv + xy(mv);
            if (cb.mvScale[3] != 0.0f) prev = screenUv(cb.worldToClipPrev, X + mv);
            const float2 delta = (prev - expected) * make_float2((float)cb.rectSize[0], (float)cb.rectSize[1]);
            set(isInScreenNearest(prev) ? make_float3(fabsf(delta.x), fabsf(delta.y), 0.0f) : make_float3(0.0f, 0.0f, 1.0f));
            break;
        }
        case 4: {
            text.print("UNITS & JITTER");
            const float2 dim = make_float2(0.5f * resourceSize.y / resourceSize.x, 0.5f);
            const float2 remapped = (cell.uv - (1.0f - dim)) / dim;
            if (remapped.x > 0.0f && remapped.y > 0.0f) {
                const float2 dimInPixels = resourceSize * 0.25f * dim;
                const float2 uv = make_float2(cb.jitter[0], cb.jitter[1]) + 0.5f;
                const bool valid = saturate(uv.x) == uv.x && saturate(uv.y) == uv.y;
                const int ax = (int)(saturate(uv.x) * dimInPixels.x), ay = (int)(saturate(uv.y) * dimInPixels.y);
                const int bx = (int)(remapped.x * dimInPixels.x), by = (int)(remapped.y * dimInPixels.y);
                float3 rgb = xyz(result);
                if (abs(ax - bx) <= 1 && abs(ay - by) <= 1 && valid) rgb = f3(0.66f);
                if (abs(ax - bx) <= 3 && abs(ay - by) <= 3 && !valid) rgb = make_float3(1.0f, 0.0f, 0.0f);
                result = f4(rgb, 1.0f);
            } else {
                const float3 shifted = X + viewZ * 0.001f;
                set(make_float3(frac(shifted.x), frac(shifted.y), frac(shifted.z)) * notInf);
            }
            break;
        }
        case 8: {
            text.print("DIFF-SPEC FRAMES");
            float f = 1.0f - saturate(historyLength / fmaxf(fmaxf(cb.diffMaxAccumulatedFrameNum, cb.specMaxAccumulatedFrameNum), 1.0f));
            if (checker && historyLength < 2.0f) f = 0.75f;
            set(colorizeZucconi(cell.uv.y > 0.95f ? 1.0f - cell.uv.x : f * notInf));
            break;
        }
        default: break;
    }
    anyStore4(p.out, px, py, f4(applyCaption(xyz(result), text.foreground), result.w));
}

// ---------------------------------------------------------------------------------------------------------------
uint32_t bytesOf(nrd::Format f) {
    switch (f) {
        case nrd::Format::R8_UNORM: return 1;
        case nrd::Format::RG8_UNORM: case nrd::Format::R16_SFLOAT: case nrd::Format::R16_UINT: return 2;
        case nrd::Format::RGBA16_SFLOAT: return 8;
        default: return 4;
    }
}
struct RelaxBinder {
    const nrdcuTexture* t;
    uint32_t n, next;
    bool ok;
    std::string* err;
    const char* id;
    template <class V> V take(nrd::Format expect) {
        V v{};
        if (next >= n) {
            ok = false;
            return v;
        }
        const nrdcuTexture& x = t[next];
        const uint32_t bpp = bytesOf(expect);
        if (x.format != (uint32_t)expect || !x.data || (x.pitchBytes % bpp) != 0 || x.pitchBytes < x.width * bpp) {
            if (ok) *err = std::string(id) + ": binding " + std::to_string(next) + " has format " + std::to_string(x.format) + " (expected " + std::to_string((uint32_t)expect) + ")";
            ok = false;
        }
        v.data = (uint8_t*)x.data;
        v.w = (int)x.width;
        v.h = (int)x.height;
        v.pitch = (int)(x.pitchBytes / bpp);
        v.fmt = x.format;
        next++;
        return v;
    }
    // any bound format the polymorphic accessors read ( validation overlay: IN_MV and OUT_VALIDATION are the application's choice )
    TexView takeView() {
        TexView v{};
        if (next >= n) {
            ok = false;
            return v;
        }
        const nrdcuTexture& x = t[next];
        const uint32_t bpp = x.format == (uint32_t)nrd::Format::RGBA32_SFLOAT ? 16u : (x.format == (uint32_t)nrd::Format::RGBA16_UNORM || x.format == (uint32_t)nrd::Format::RGBA16_SNORM) ? 8u
                                                                                      : (x.format == (uint32_t)nrd::Format::R16_UNORM ? 2u : bytesOf((nrd::Format)x.format));
        if (!x.data || (x.pitchBytes % bpp) != 0 || x.pitchBytes < x.width * bpp) {
            if (ok) *err = std::string(id) + ": binding " + std::to_string(next) + " has a pitch that does not fit its format";
            ok = false;
        }
        v.data = (uint8_t*)x.data;
        v.w = (int)x.width;
        v.h = (int)x.height;
        v.pitch = (int)(x.pitchBytes / bpp);
        v.fmt = x.format;
        next++;
        return v;
    }
    // optional single-channel guides ( IN_VIEWZ as a dummy when disabled, otherwise the application's texture: any size, see bindGuide )
    TexAnyX takeGuide() {
        TexAnyX v{};
        if (next >= n) {
            ok = false;
            return v;
        }
        const nrdcuTexture& x = t[next];
        if (!bindGuide(x.format, x.data, x.width, x.height, x.pitchBytes, v)) {
            if (ok) *err = std::string(id) + ": binding " + std::to_string(next) + " (single-channel guide) has unsupported format " + std::to_string(x.format);
            ok = false;
        }
        next++;
        return v;
    }
};

}  // namespace

// Dispatch by shader identifier (called by the executor). Returns an nrd::Result; `err` explains a failure.
uint32_t dispatchRelax(const PipelineKey& key, const void* constants, uint32_t constantsSize, const nrdcuTexture* tex, uint32_t n, Rows rows, cudaStream_t stream, std::string& err) {
    using nrd::Format;
    using nrd::Result;
    const char* const id = key.id;
    if (constantsSize != sizeof(RelaxConstants) || !constants) {
        err = std::string(id) + ": expected " + std::to_string(sizeof(RelaxConstants)) + " constant bytes";
        return (uint32_t)Result::INVALID_ARGUMENT;
    }
    RelaxConstants cb;
    memcpy(&cb, constants, sizeof(cb));
    if (cb.rectOrigin[0] || cb.rectOrigin[1]) {   // NRD_SUPPORTS_VIEWPORT_OFFSET = 0; dynamic resolution ( rectSize < resourceSize ) itself is supported
        err = std::string(id) + ": rectOrigin must be 0";
        return (uint32_t)Result::INVALID_ARGUMENT;
    }
    if ((cb.diffCheckerboard == 2) != (cb.specCheckerboard == 2)) {
        err = std::string(id) + ": inconsistent checkerboard constants";
        return (uint32_t)Result::INVALID_ARGUMENT;
    }
    const bool checkerboarded = cb.diffCheckerboard != 2;
    RelaxBinder b{tex, n, 0, true, &err, id};
    auto bad = [&](uint32_t expected) {
        if (b.ok && b.next == expected && n == expected) return false;
        if (err.empty()) err = std::string(id) + ": wrong number of textures";
        return true;
    };
    const Format F16 = Format::RGBA16_SFLOAT, R8 = Format::R8_UNORM, R32 = Format::R32_SFLOAT, NR = Format::R10_G10_B10_A2_UNORM;
    const dim3 block(BLOCK_W, BLOCK_H);
    // rows [ begin, end ) of the rect ( multi-GPU strips, nrdcuDenoiseRows ): every pass covers only the CTA rows of the strip ( begin is a multiple of 16 )
    const RowGrid rg = rowGrid(rows, (int)cb.rectSize[1], BLOCK_H);
    const int ctaY0 = rg.ctaY0;
    const dim3 pixelGrid((cb.rectSize[0] + BLOCK_W - 1) / BLOCK_W, rg.count);
    if (rg.count == 0 && key.pass != RELAX_VALIDATION) return (uint32_t)Result::SUCCESS;
    // "|NRD_SIGNAL=<DIFF|SPEC|BOTH>|NRD_MODE=<SH|RADIANCE>": the six RELAX denoisers run the same kernels; RADIANCE has no SH1 textures, a single-lobe
    // denoiser binds only its own lobe ( the parameter block keeps the two-lobe layout, the other lobe's views stay empty and are dead code in the kernel )
    const bool sh = key.mode == 1;
    const int signal = key.signal == 1 ? SIGNAL_DIFF : (key.signal == 2 ? SIGNAL_SPEC : SIGNAL_BOTH);
    const bool hasDiff = (signal & SIGNAL_DIFF) != 0, hasSpec = (signal & SIGNAL_SPEC) != 0;
    const uint32_t lobes = (hasDiff ? 1u : 0u) + (hasSpec ? 1u : 0u);
    auto takeS = [&](TexRGBA16F& t) { t = hasSpec ? b.take<TexRGBA16F>(F16) : TexRGBA16F(); };
    auto takeD = [&](TexRGBA16F& t) { t = hasDiff ? b.take<TexRGBA16F>(F16) : TexRGBA16F(); };
    auto takeShS = [&](TexRGBA16F& t) { t = (sh && hasSpec) ? b.take<TexRGBA16F>(F16) : TexRGBA16F(); };
    auto takeShD = [&](TexRGBA16F& t) { t = (sh && hasDiff) ? b.take<TexRGBA16F>(F16) : TexRGBA16F(); };
    const uint32_t shOn = sh ? 1u : 0u;

    if (key.pass == RELAX_VALIDATION) {
        RelaxValidationParams p;
        p.normalRoughness = b.take<TexNR>(NR);
        p.viewZ = b.take<TexR32F>(R32);
        p.mv = b.takeView();
        p.historyLength = b.take<TexR8>(R8);
        p.out = b.takeView();
        if (bad(5)) return (uint32_t)Result::INVALID_ARGUMENT;
        launchK(relaxValidationKernel, dim3((p.out.w + 31) / 32, (p.out.h + 7) / 8), 256, 0, stream, cb, p);
    } else if (key.pass == RELAX_CLASSIFY_TILES) {
        RelaxClassifyParams p;
        p.viewZ = b.take<TexR32F>(R32);
        p.outTiles = b.take<TexR8>(R8);
        if (bad(2)) return (uint32_t)Result::INVALID_ARGUMENT;
        const RowGrid tg = rowGrid(rows, (int)cb.rectSize[1], 16);
        const int tilesW = (cb.rectSize[0] + 15) / 16;
        launchK(relaxClassifyTilesKernel, dim3((tilesW + 7) / 8, tg.count), 256, 0, stream, cb, p, tg.ctaY0, tilesW);
    } else if (key.pass == RELAX_PREPASS) {
        RelaxPrePassParams p;
        p.tiles = b.take<TexR8>(R8);
        p.normalRoughness = b.take<TexNR>(NR);
        p.viewZ = b.take<TexR32F>(R32);
        takeS(p.spec);
        takeD(p.diff);
        takeShS(p.specSh);
        takeShD(p.diffSh);
        takeS(p.outSpec);
        takeD(p.outDiff);
        takeShS(p.outSpecSh);
        takeShD(p.outDiffSh);
        if (bad(3 + (2 + 2 * shOn) * lobes)) return (uint32_t)Result::INVALID_ARGUMENT;
        if (checkerboarded) {
            if (sh) withSignal(signal, [&](auto sig_) { launchK(relaxPrePassKernel<true, true, decltype(sig_)::value>, pixelGrid, block, 0, stream, cb, lobeView(p, sig_), ctaY0); }); else withSignal(signal, [&](auto sig_) { launchK(relaxPrePassKernel<false, true, decltype(sig_)::value>, pixelGrid, block, 0, stream, cb, lobeView(p, sig_), ctaY0); });
        } else {
            if (sh) withSignal(signal, [&](auto sig_) { launchK(relaxPrePassKernel<true, false, decltype(sig_)::value>, pixelGrid, block, 0, stream, cb, lobeView(p, sig_), ctaY0); }); else withSignal(signal, [&](auto sig_) { launchK(relaxPrePassKernel<false, false, decltype(sig_)::value>, pixelGrid, block, 0, stream, cb, lobeView(p, sig_), ctaY0); });
        }
    } else if (key.pass == RELAX_TEMPORAL_ACCUMULATION) {
        RelaxTaParams p;
        p.tiles = b.take<TexR8>(R8);
        p.mv = b.take<TexRGBA16F>(F16);
        p.normalRoughness = b.take<TexNR>(NR);
        p.viewZ = b.take<TexR32F>(R32);
        p.mixDummy = b.takeGuide();
        p.prevNormalRoughness = b.take<TexRGBA8>(Format::RGBA8_UNORM);
        p.prevViewZ = b.take<TexR32F>(R32);
        p.prevHistoryLength = b.take<TexR8>(R8);
        p.prevMaterialID = b.take<TexR8>(R8);
        takeS(p.spec);
        takeD(p.diff);
        takeS(p.historySpecFast);
        takeD(p.historyDiffFast);
        takeS(p.historySpec);
        takeD(p.historyDiff);
        if (hasSpec) p.prevSpecHitDist = b.take<TexR16F>(Format::R16_SFLOAT);
        if (hasSpec) p.specConfDummy = b.takeGuide();
        if (hasDiff) p.diffConfDummy = b.takeGuide();
        takeShS(p.specSh);
        takeShD(p.diffSh);
        takeShS(p.historySpecShFast);
        takeShD(p.historyDiffShFast);
        takeShS(p.historySpecSh);
        takeShD(p.historyDiffSh);
        p.outHistoryLength = b.take<TexR8>(R8);
        takeS(p.outSpec);
        takeD(p.outDiff);
        takeS(p.outSpecFast);
        takeD(p.outDiffFast);
        if (hasSpec) p.outSpecHitDist = b.take<TexR16F>(Format::R16_SFLOAT);
        if (hasSpec) p.outSpecReprojectionConfidence = b.take<TexR8>(R8);
        takeShS(p.outSpecSh);
        takeShD(p.outDiffSh);
        takeShS(p.outSpecShFast);
        takeShD(p.outDiffShFast);
        if (bad(10 + (6 + 5 * shOn) * lobes + (hasSpec ? 3 : 0) + 0 * shOn)) return (uint32_t)Result::INVALID_ARGUMENT;
        if (checkerboarded || cb.hasHistoryConfidence || cb.hasDisocclusionThresholdMix) {
            if (sh) withSignal(signal, [&](auto sig_) { launchK(relaxTemporalAccumulationKernel<true, true, decltype(sig_)::value>, pixelGrid, block, 0, stream, cb, lobeView(p, sig_), ctaY0); }); else withSignal(signal, [&](auto sig_) { launchK(relaxTemporalAccumulationKernel<false, true, decltype(sig_)::value>, pixelGrid, block, 0, stream, cb, lobeView(p, sig_), ctaY0); });
        } else {
            if (sh) withSignal(signal, [&](auto sig_) { launchK(relaxTemporalAccumulationKernel<true, false, decltype(sig_)::value>, pixelGrid, block, 0, stream, cb, lobeView(p, sig_), ctaY0); }); else withSignal(signal, [&](auto sig_) { launchK(relaxTemporalAccumulationKernel<false, false, decltype(sig_)::value>, pixelGrid, block, 0, stream, cb, lobeView(p, sig_), ctaY0); });
        }
    } else if (key.pass == RELAX_HISTORY_FIX) {
        RelaxHistoryFixParams p;
        p.tiles = b.take<TexR8>(R8);
        p.historyLength = b.take<TexR8>(R8);
        p.normalRoughness = b.take<TexNR>(NR);
        p.viewZ = b.take<TexR32F>(R32);
        takeS(p.spec);
        takeD(p.diff);
        takeShS(p.specSh);
        takeShD(p.diffSh);
        takeS(p.outSpec);
        takeD(p.outDiff);
        takeShS(p.outSpecSh);
        takeShD(p.outDiffSh);
        if (bad(4 + (2 + 2 * shOn) * lobes)) return (uint32_t)Result::INVALID_ARGUMENT;
        if (sh) withSignal(signal, [&](auto sig_) { launchK(relaxHistoryFixKernel<true, decltype(sig_)::value>, pixelGrid, block, 0, stream, cb, lobeView(p, sig_), ctaY0); }); else withSignal(signal, [&](auto sig_) { launchK(relaxHistoryFixKernel<false, decltype(sig_)::value>, pixelGrid, block, 0, stream, cb, lobeView(p, sig_), ctaY0); });
    } else if (key.pass == RELAX_HISTORY_CLAMPING) {
        RelaxHistoryClampingParams p;
        p.tiles = b.take<TexR8>(R8);
        p.viewZ = b.take<TexR32F>(R32);
        p.historyLength = b.take<TexR8>(R8);
        TexRGBA16F* ins[6] = {&p.specNoisy, &p.diffNoisy, &p.spec, &p.diff, &p.specFast, &p.diffFast};
        for (int i = 0; i < 6; i++) (i & 1) ? takeD(*ins[i]) : takeS(*ins[i]);
        TexRGBA16F* insSh[4] = {&p.specSh, &p.diffSh, &p.specShFast, &p.diffShFast};
        for (int i = 0; i < 4; i++) (i & 1) ? takeShD(*insSh[i]) : takeShS(*insSh[i]);
        p.outHistoryLength = b.take<TexR8>(R8);
        TexRGBA16F* outs[4] = {&p.outSpec, &p.outDiff, &p.outSpecFast, &p.outDiffFast};
        for (int i = 0; i < 4; i++) (i & 1) ? takeD(*outs[i]) : takeS(*outs[i]);
        TexRGBA16F* outsSh[4] = {&p.outSpecSh, &p.outDiffSh, &p.outSpecShFast, &p.outDiffShFast};
        for (int i = 0; i < 4; i++) (i & 1) ? takeShD(*outsSh[i]) : takeShS(*outsSh[i]);
        if (bad(4 + (5 + 4 * shOn) * lobes)) return (uint32_t)Result::INVALID_ARGUMENT;
        // 32x16 pixels per CTA ( two rows per thread ); strips begin on multiples of 16 rows
        const RowGrid rgHc = rowGrid(rows, (int)cb.rectSize[1], HC_PIXELS_H);
        const dim3 gridHc(pixelGrid.x, rgHc.count);
        const int rowEnd = std::min(rows.end, (int)cb.rectSize[1]);
        if (sh) withSignal(signal, [&](auto sig_) { launchK(relaxHistoryClampingKernel<true, decltype(sig_)::value>, gridHc, block, 0, stream, cb, lobeView(p, sig_), rgHc.ctaY0, rowEnd); }); else withSignal(signal, [&](auto sig_) { launchK(relaxHistoryClampingKernel<false, decltype(sig_)::value>, gridHc, block, 0, stream, cb, lobeView(p, sig_), rgHc.ctaY0, rowEnd); });
    } else if (key.pass == RELAX_HITDIST_RECONSTRUCTION) {
        RelaxHitDistReconstructionParams p;
        p.tiles = b.take<TexR8>(R8);
        p.normalRoughness = b.take<TexNR>(NR);
        p.viewZ = b.take<TexR32F>(R32);
        takeS(p.spec);
        takeD(p.diff);
        takeS(p.outSpec);
        takeD(p.outDiff);
        if (bad(3 + 2 * lobes)) return (uint32_t)Result::INVALID_ARGUMENT;
        if (key.mode5x5)
            launchK(relaxHitDistReconstructionKernel<2>, pixelGrid, block, 0, stream, cb, p, ctaY0);
        else
            launchK(relaxHitDistReconstructionKernel<1>, pixelGrid, block, 0, stream, cb, p, ctaY0);
    } else if (key.pass == RELAX_SPLIT_SCREEN) {
        RelaxSplitScreenParams p;
        p.viewZ = b.take<TexR32F>(R32);
        takeD(p.diff);
        takeS(p.spec);
        takeShD(p.diffSh);
        takeShS(p.specSh);
        takeD(p.outDiff);
        takeS(p.outSpec);
        takeShD(p.outDiffSh);
        takeShS(p.outSpecSh);
        if (bad(1 + (2 + 2 * shOn) * lobes)) return (uint32_t)Result::INVALID_ARGUMENT;
        if (sh) launchK(relaxSplitScreenKernel<true>, pixelGrid, block, 0, stream, cb, p, ctaY0); else launchK(relaxSplitScreenKernel<false>, pixelGrid, block, 0, stream, cb, p, ctaY0);
    } else if (key.pass == RELAX_COPY) {
        RelaxCopyParams p;
        takeS(p.spec);
        takeD(p.diff);
        takeS(p.outSpec);
        takeD(p.outDiff);
        if (bad(2 * lobes)) return (uint32_t)Result::INVALID_ARGUMENT;
        withSignal(signal, [&](auto sig_) { launchK(relaxCopyKernel<decltype(sig_)::value>, pixelGrid, block, 0, stream, lobeView(p, sig_), ctaY0); });
    } else if (key.pass == RELAX_ANTI_FIREFLY) {
        RelaxAntiFireflyParams p;
        p.tiles = b.take<TexR8>(R8);
        p.normalRoughness = b.take<TexNR>(NR);
        p.viewZ = b.take<TexR32F>(R32);
        takeS(p.spec);
        takeD(p.diff);
        takeS(p.outSpec);
        takeD(p.outDiff);
        if (bad(3 + 2 * lobes)) return (uint32_t)Result::INVALID_ARGUMENT;
        withSignal(signal, [&](auto sig_) { launchK(relaxAntiFireflyKernel<decltype(sig_)::value>, pixelGrid, block, 0, stream, cb, lobeView(p, sig_), ctaY0); });
    } else if (key.pass == RELAX_ATROUS_SMEM || key.pass == RELAX_ATROUS) {
        const bool smem = key.pass == RELAX_ATROUS_SMEM;
        RelaxAtrousParams p = {};
        p.tiles = b.take<TexR8>(R8);
        p.historyLength = b.take<TexR8>(R8);
        p.normalRoughness = b.take<TexNR>(NR);
        p.viewZ = b.take<TexR32F>(R32);
        takeS(p.spec);
        takeD(p.diff);
        if (hasSpec) p.specReprojectionConfidence = b.take<TexR8>(R8);
        if (hasSpec) p.specConfDummy = b.takeGuide();
        if (hasDiff) p.diffConfDummy = b.takeGuide();
        takeShS(p.specSh);
        takeShD(p.diffSh);
        takeS(p.outSpec);
        takeD(p.outDiff);
        if (smem) {
            p.outNormalRoughness = b.take<TexRGBA8>(Format::RGBA8_UNORM);
            p.outMaterialID = b.take<TexR8>(R8);
            p.outViewZ = b.take<TexR32F>(R32);
        }
        takeShS(p.outSpecSh);
        takeShD(p.outDiffSh);
        if (bad(4 + (smem ? 3 : 0) + (3 + 2 * shOn) * lobes + (hasSpec ? 1 : 0))) return (uint32_t)Result::INVALID_ARGUMENT;
        if (smem) {
            auto launchSmem = [&](auto sh_, auto sig_) {
                constexpr bool SH_ = decltype(sh_)::value;
                constexpr int SIG_ = decltype(sig_)::value;
                if (cb.roughnessEdgeStoppingEnabled) launchK(relaxAtrousSmemKernel<SH_, SIG_, true>, pixelGrid, block, 0, stream, cb, lobeView(p, sig_), ctaY0);
                else launchK(relaxAtrousSmemKernel<SH_, SIG_, false>, pixelGrid, block, 0, stream, cb, lobeView(p, sig_), ctaY0);
            };
            if (sh) withSignal(signal, [&](auto sig_) { launchSmem(std::true_type(), sig_); }); else withSignal(signal, [&](auto sig_) { launchSmem(std::false_type(), sig_); });
        } else {
            // strides 2 and 4: the neighbourhood is staged in shared memory ( relaxAtrousTiledKernel ); larger strides gather
            auto launchAtrous = [&](auto sh_, auto sig_, auto res_) {
                constexpr bool SH_ = decltype(sh_)::value, RES_ = decltype(res_)::value;
                constexpr int SIG_ = decltype(sig_)::value;
                if (cb.stepSize == 2 || cb.stepSize == 4) {
                    // the raw planes through TMA when every bound plane can be described by a tensor map ( NRD_B200_RELAX_TMA=0: A/B switch of the benchmarks and of tests/test_relax_tma_gpu.py )
                    constexpr bool HAS_SPEC_ = (SIG_ & SIGNAL_SPEC) != 0, HAS_DIFF_ = (SIG_ & SIGNAL_DIFF) != 0;
                    const char* tmaEnv = getenv("NRD_B200_RELAX_TMA");   // read per dispatch: tests flip it inside one process
                    const bool tmaWanted = !(tmaEnv && tmaEnv[0] == '0');
                    const int step = (int)cb.stepSize, boxW = BLOCK_W + 2 * step, boxH = BLOCK_H + 2 * step;
                    RelaxAtrousTma tma;
                    memset(&tma, 0, sizeof(tma));
                    bool useTma = tmaWanted;
                    if (useTma && HAS_SPEC_) useTma = encodeRgba16fBox(tma.spec, p.spec, boxW, boxH) && (!SH_ || encodeRgba16fBox(tma.specSh, p.specSh, boxW, boxH));
                    if (useTma && HAS_DIFF_) useTma = encodeRgba16fBox(tma.diff, p.diff, boxW, boxH) && (!SH_ || encodeRgba16fBox(tma.diffSh, p.diffSh, boxW, boxH));
                    if (step == 2) {
                        if (useTma) launchK(relaxAtrousTiledKernel<SH_, SIG_, 2, true, RES_>, pixelGrid, block, 0, stream, cb, lobeView(p, sig_), ctaY0, tma);
                        else launchK(relaxAtrousTiledKernel<SH_, SIG_, 2, false, RES_>, pixelGrid, block, 0, stream, cb, lobeView(p, sig_), ctaY0, tma);
                    } else {
                        if (useTma) launchK(relaxAtrousTiledKernel<SH_, SIG_, 4, true, RES_>, pixelGrid, block, 0, stream, cb, lobeView(p, sig_), ctaY0, tma);
                        else launchK(relaxAtrousTiledKernel<SH_, SIG_, 4, false, RES_>, pixelGrid, block, 0, stream, cb, lobeView(p, sig_), ctaY0, tma);
                    }
                }
                else launchK(relaxAtrousKernel<SH_, SIG_, RES_>, pixelGrid, block, 0, stream, cb, lobeView(p, sig_), ctaY0);
            };
            auto launchAtrousSh = [&](auto sh_, auto sig_) {
                if (cb.roughnessEdgeStoppingEnabled) launchAtrous(sh_, sig_, std::true_type()); else launchAtrous(sh_, sig_, std::false_type());
            };
            if (sh) withSignal(signal, [&](auto sig_) { launchAtrousSh(std::true_type(), sig_); }); else withSignal(signal, [&](auto sig_) { launchAtrousSh(std::false_type(), sig_); });
        }
    } else {
        err = std::string("no CUDA kernel for shader '") + id + "'";
        return (uint32_t)Result::UNSUPPORTED;
    }
    return (uint32_t)Result::SUCCESS;
}

}  // namespace nrdk
