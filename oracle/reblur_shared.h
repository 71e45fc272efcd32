// TEST INFRASTRUCTURE — CPU oracle (see hlsl_like.h).
// REBLUR constant buffer + the helpers of REBLUR_Config.hlsli / REBLUR_Common.hlsli / Common.hlsli that read it.
// cb layout: External/NRD/Shaders/REBLUR_Config.hlsli:115-192 (864 bytes, HLSL packing; float4x4 column-major).
#pragma once
#include "nrd_shared.h"

namespace orc {

struct uint2 { uint32_t x, y; };

struct ReblurCB {
    float4x4 gWorldToClip, gViewToClip, gViewToWorld, gWorldToViewPrev, gWorldToClipPrev, gWorldPrevToWorld;
    float4 gRotatorPre, gRotator, gRotatorPost, gFrustum, gFrustumPrev, gCameraDelta, gHitDistSettings, gViewVectorWorld, gViewVectorWorldPrev, gMvScale,
        gConvergenceSettings;
    float2 gAntilagSettings, gResourceSize, gResourceSizeInv, gResourceSizeInvPrev, gRectSize, gRectSizeInv, gRectSizePrev, gResolutionScale,
        gResolutionScalePrev, gRectOffset, gJitter;
    uint2 gPrintfAt, gRectOrigin;
    int2 gRectSizeMinusOne;
    float gDisocclusionThreshold, gDisocclusionThresholdAlternate, gCameraAttachedReflectionMaterialID, gStrandMaterialID, gStrandThickness,
        gStabilizationStrength, gDebug, gOrthoMode, gUnproject, gDenoisingRange, gPlaneDistSensitivity, gFramerateScale, gMinBlurRadius, gMaxBlurRadius,
        gDiffPrepassBlurRadius, gSpecPrepassBlurRadius, gMaxAccumulatedFrameNum, gMaxFastAccumulatedFrameNum, gAntiFirefly, gLobeAngleFraction,
        gRoughnessFraction, gHistoryFixFrameNum, gHistoryFixBasePixelStride, gHistoryFixAlternatePixelStride, gHistoryFixAlternatePixelStrideMaterialID,
        gFastHistoryClampingSigmaScale, gMinRectDimMulUnproject, gUsePrepassNotOnlyForSpecularMotionEstimation, gSplitScreen, gSplitScreenPrev,
        gCheckerboardResolveAccumSpeed, gViewZScale, gFireflySuppressorMinRelativeScale, gMinHitDistanceWeight, gDiffMinMaterial, gSpecMinMaterial,
        gResponsiveAccumulationInvRoughnessThreshold;
    uint32_t gResponsiveAccumulationMinAccumulatedFrameNum, gHasHistoryConfidence, gHasDisocclusionThresholdMix, gDiffCheckerboard, gSpecCheckerboard,
        gFrameIndex, gIsRectChanged, gResetHistory, gReturnHistoryLengthInsteadOfOcclusion;
    uint32_t _pad[2];
};
static_assert(sizeof(ReblurCB) == 864, "REBLUR cbuffer is 864 bytes");

// REBLUR_Config.hlsli settings
static const int REBLUR_MAX_ACCUM_FRAME_NUM = 63;   // 6 bits
static const int REBLUR_MAX_MATERIALID_NUM = 15;    // 4 bits
static const float REBLUR_INVALID = -32768.0f;
static const float REBLUR_FIREFLY_SUPPRESSOR_MAX_RELATIVE_INTENSITY = 38.0f;
static const float REBLUR_FIREFLY_SUPPRESSOR_RADIUS_SCALE = 0.1f;
static const float REBLUR_FIREFLY_SUPPRESSOR_FAST_RELATIVE_INTENSITY = 4.0f;
static const float REBLUR_ANTI_FIREFLY_SIGMA_SCALE = 2.0f;
static const float REBLUR_ROUGHNESS_SENSITIVITY_IN_TA = NRD_ROUGHNESS_SENSITIVITY * 0.3f;
static const float REBLUR_MAX_PERCENT_OF_LOBE_VOLUME_FOR_PRE_PASS = 0.3f;

// Everything below closes over one frame's constants, the way the shaders see the cbuffer as globals.
struct ReblurCtx {
    const ReblurCB& cb;
    const int signal;  // NRD_SIGNAL: 1 = DIFF, 2 = SPEC, 3 = BOTH (NRD.hlsli:338-339)
    explicit ReblurCtx(const ReblurCB& c, int sig = 3) : cb(c), signal(sig) {}
    bool hasDiff() const { return (signal & 1) != 0; }
    bool hasSpec() const { return (signal & 2) != 0; }

    float UnpackViewZ(float z) const { return std::fabs(z * cb.gViewZScale); }            // common:261
    bool IsInDenoisingRange(float z) const { return z < cb.gDenoisingRange; }             // common:262 (false for NaN)
    float ApplyGeometryWeightLast(float w, float z, float NoX, float2 p) const {          // common:567
        w *= ComputeWeight(NoX, p.x, p.y);
        return !IsInDenoisingRange(z) ? 0.0f : w;
    }

    // ---- REBLUR_Common.hlsli ----
    uint32_t PackInternalData(float diffAccumSpeed, float specAccumSpeed, float materialID) const {  // :13
        diffAccumSpeed = min(diffAccumSpeed + 1.0f, cb.gMaxAccumulatedFrameNum);
        specAccumSpeed = min(specAccumSpeed + 1.0f, cb.gMaxAccumulatedFrameNum);
        float3 t;
        t.x = hlsl_round(diffAccumSpeed) / REBLUR_MAX_ACCUM_FRAME_NUM;
        t.y = hlsl_round(specAccumSpeed) / REBLUR_MAX_ACCUM_FRAME_NUM;
        t.z = materialID / REBLUR_MAX_MATERIALID_NUM;
        return Packing::RgbaToUint(float4(t.x, t.y, t.z, t.z), 6, 6, 4, 0);
    }
    static float3 UnpackInternalData(uint32_t p) {  // :29
        float4 t = Packing::UintToRgba(p, 6, 6, 4, 0);
        return float3(hlsl_round(t.x * REBLUR_MAX_ACCUM_FRAME_NUM), hlsl_round(t.y * REBLUR_MAX_ACCUM_FRAME_NUM), t.z * REBLUR_MAX_MATERIALID_NUM);
    }
    // `signal` = NRD_SIGNAL of the permutation: a specular-only denoiser keeps its one value in .x of an R8_UNORM texture (:48-51, :58-61),
    // a diffuse-only one stores data2 in 8 bits with the CatRom flag in bit 4 instead of 15 (:69-73)
    static float2 PackData1(float diffAccumSpeed, float specAccumSpeed, int signal = 3) {  // :42
        float2 r = float2(saturate(hlsl_round(diffAccumSpeed) / REBLUR_MAX_ACCUM_FRAME_NUM), saturate(hlsl_round(specAccumSpeed) / REBLUR_MAX_ACCUM_FRAME_NUM));
        if (!(signal & 1)) r.x = r.y;
        return r;
    }
    static float2 UnpackData1(float2 p, int signal = 3) {  // :56
        if (!(signal & 1)) p.y = p.x;
        return float2(hlsl_round(p.x * REBLUR_MAX_ACCUM_FRAME_NUM), hlsl_round(p.y * REBLUR_MAX_ACCUM_FRAME_NUM));
    }
    static uint32_t PackData2(float fbits, float curvature, float virtualHistoryAmount, bool smbAllowCatRom, int signal = 3) {  // :75
        const uint32_t smbAllowCatRomBit = (signal & 2) ? 15u : 4u;
        uint32_t p = (uint32_t)(fbits + 0.5f);
        p |= (uint32_t)(saturate(virtualHistoryAmount) * 127.0f + 0.5f) << 8;
        p |= smbAllowCatRom ? (1u << smbAllowCatRomBit) : 0u;
        p |= (uint32_t)f32tof16(curvature) << 16;
        return p;
    }
    static float2 UnpackData2(uint32_t p, uint32_t& bits, bool& smbAllowCatRom, int signal = 3) {  // :92
        const uint32_t smbAllowCatRomBit = (signal & 2) ? 15u : 4u;
        bits = p & 0xFFu;
        smbAllowCatRom = (p & (1u << smbAllowCatRomBit)) != 0;
        return float2(float((p >> 8) & 127u) / 127.0f, f16tof32(p >> 16));
    }
    float3 GetViewVector(float3 X, bool isViewSpace = false) const {  // :105
        return cb.gOrthoMode == 0.0f ? normalize(-X) : (isViewSpace ? float3(0, 0, -1) : cb.gViewVectorWorld.xyz());
    }
    float3 GetViewVectorPrev(float3 Xprev, float3 cameraDelta) const {  // :110
        return cb.gOrthoMode == 0.0f ? normalize(cameraDelta - Xprev) : cb.gViewVectorWorldPrev.xyz();
    }
    float GetMinAllowedLimitForHitDistNonLinearAccumSpeed(float roughness) const {  // :115
        float frameNum = 0.5f * GetSpecMagicCurve(roughness) * cb.gMaxAccumulatedFrameNum;
        return 1.0f / (1.0f + frameNum);
    }
    float RemapRoughnessToResponsiveFactor(float roughness) const {  // :125
        return Math::SmoothStep01(max(roughness, 1e-3f) * cb.gResponsiveAccumulationInvRoughnessThreshold);
    }
    static float GetLumaScale(float currLuma, float newLuma) { return (newLuma + NRD_EPS) / (currLuma + NRD_EPS); }  // :138
    float4 MixHistoryAndCurrent(float4 history, float4 current, float f, float roughness = 1.0f) const {  // :199 (RADIANCE mode)
        float4 r;
        r.x = lerp(history.x, current.x, f);
        r.y = lerp(history.y, current.y, f);
        r.z = lerp(history.z, current.z, f);
        r.w = lerp(history.w, current.w, max(f, GetMinAllowedLimitForHitDistNonLinearAccumSpeed(roughness)));
        return r;
    }
    static float GetLuma(float4 c) { return c.x; }  // :211, REBLUR_USE_YCOCG = 1
    static float4 ChangeLuma(float4 c, float newLuma) {  // :220
        float s = GetLumaScale(GetLuma(c), newLuma);
        return float4(c.x * s, c.y * s, c.z * s, c.w);
    }
    static float4 ClampNegativeToZero(float4 c) {  // :227
        float3 rgb = _NRD_LinearToYCoCg(_NRD_YCoCgToLinear(c.xyz()));
        return float4(rgb, saturate(c.w));
    }
    float ComputeAntilag(float h, float a, float sigma, float accumSpeed) const {  // :243 (REBLUR_ANTILAG_MODE = 2)
        float s = sigma * cb.gAntilagSettings.x;
        float magic = cb.gAntilagSettings.y * cb.gFramerateScale * cb.gFramerateScale;
        float hc = Color::Clamp(a, s, h);
        float d = std::fabs(h - hc) / (max(h, hc) + NRD_EPS);
        return 1.0f / (1.0f + d * accumSpeed / magic);
    }
    static void GetKernelBasis(float3 D, float3 N, float3& T, float3& B) {  // :275
        Geometry::Basis basis = Geometry::GetBasis(N);
        T = basis.T;
        B = basis.B;
        if (std::fabs(dot(D, N)) < 0.999f) {
            float3 R = reflect(-D, N);
            T = normalize(cross(N, R));
            B = cross(R, T);
        }
    }
    float GetNonLinearAccumSpeed(float accumSpeed, float maxAccumSpeed, float confidence, bool hasData) const {  // :294
        float n = max(1.0f - confidence, 1.0f / (1.0f + min(accumSpeed, maxAccumSpeed)));
        if (!hasData) n *= lerp(1.0f - cb.gCheckerboardResolveAccumSpeed, 1.0f, n);
        return n;
    }
    float GetAdvancedNonLinearAccumSpeed(float accumSpeed) const {  // :308
        float f = saturate(accumSpeed / (1.0f + cb.gMaxAccumulatedFrameNum * cb.gConvergenceSettings.z));
        float e = cb.gConvergenceSettings.x * lerp(cb.gConvergenceSettings.y, 1.0f, f);
        return 1.0f / (1.0f + e * accumSpeed);
    }
    float2 GetTemporalAccumulationParams(float isInScreenMulFootprintQuality, float accumSpeed, float antilag) const {  // :317
        float w = isInScreenMulFootprintQuality;
        w *= 1.0f - GetAdvancedNonLinearAccumSpeed(accumSpeed);
        w *= antilag;
        return float2(w, 1.0f + 3.0f * cb.gFramerateScale * w);
    }
};

// Common.hlsli:604-658 — 12-tap Catmull-Rom without corners (5 bilinear fetches) with fallback to a custom-weighted
// bilinear tap set. `samplePos` is in texels, `invResourceSize` converts texels to uv.
struct HistoryFilter {
    float4 w;            // weights of taps 0..3
    float w4;            // weight of tap 4
    float sum;
    float2 uv[5];
    int2 bilinearOrigin;
    float4 bilinearCustomWeights;
    HistoryFilter(float2 samplePos, float2 invResourceSize, float4 customWeights, bool useBicubic) {
        const float S = NRD_CATROM_SHARPNESS;
        float2 centerPos = floor(samplePos - 0.5f) + 0.5f;
        float2 f = saturate(samplePos - centerPos);
        float2 w0 = f * (f * (-S * f + 2.0f * S) - S);
        float2 w1 = f * (f * ((2.0f - S) * f - (3.0f - S))) + 1.0f;
        float2 w2 = f * (f * (-(2.0f - S) * f + (3.0f - 2.0f * S)) + S);
        float2 w3 = f * (f * (S * f - S));
        float2 w12 = w1 + w2;
        float2 tc = w2 / w12;
        w.x = w12.x * w0.y;
        w.y = w0.x * w12.y;
        w.z = w12.x * w12.y;
        w.w = w3.x * w12.y;
        w4 = w12.x * w3.y;
        w = useBicubic ? w : customWeights;
        w4 = useBicubic ? w4 : 0.0f;
        sum = dot(w, float4(1.0f)) + w4;
        if (useBicubic) {
            uv[0] = centerPos + float2(tc.x, -1.0f);
            uv[1] = centerPos + float2(-1.0f, tc.y);
            uv[2] = centerPos + float2(tc.x, tc.y);
            uv[3] = centerPos + float2(2.0f, tc.y);
            uv[4] = centerPos + float2(tc.x, 2.0f);
        } else {
            uv[0] = centerPos + float2(0, 0);
            uv[1] = centerPos + float2(1, 0);
            uv[2] = centerPos + float2(0, 1);
            uv[3] = centerPos + float2(1, 1);
            uv[4] = centerPos + f;
        }
        for (int i = 0; i < 5; i++) uv[i] = uv[i] * invResourceSize;
        bilinearOrigin = int2((int)centerPos.x, (int)centerPos.y);  // int3( centerPos, 0 ): truncation of x.5 >= 0 ... or negative .5
        bilinearCustomWeights = customWeights;
    }
    // CatRom (or fallback) fetch through the linear sampler
    float4 color(const Tex& tex) const {
        float4 c = tex.sampleLinear(uv[0]) * w.x;
        c += tex.sampleLinear(uv[1]) * w.y;
        c += tex.sampleLinear(uv[2]) * w.z;
        c += tex.sampleLinear(uv[3]) * w.w;
        c += tex.sampleLinear(uv[4]) * w4;
        return sum < 0.0001f ? float4(0.0f) : c / sum;
    }
    // bilinear with custom weights through Load (OOB -> 0)
    float4 bilinear(const Tex& tex) const {
        float4 c = tex.load(bilinearOrigin.x, bilinearOrigin.y) * bilinearCustomWeights.x;
        c += tex.load(bilinearOrigin.x + 1, bilinearOrigin.y) * bilinearCustomWeights.y;
        c += tex.load(bilinearOrigin.x, bilinearOrigin.y + 1) * bilinearCustomWeights.z;
        c += tex.load(bilinearOrigin.x + 1, bilinearOrigin.y + 1) * bilinearCustomWeights.w;
        float s = dot(bilinearCustomWeights, float4(1.0f));
        return s < 0.0001f ? float4(0.0f) : c / s;
    }
};

}  // namespace orc
