// TEST INFRASTRUCTURE — CPU oracle (see hlsl_like.h). NOT part of the product.
// PINNED: bit-identical, dispatch by dispatch, to the reference's own shaders compiled as C++
// (oracle/_ref/libnrd_refshaders.so, tests/test_oracle_vs_reference_shaders.py).
//
// RELAX_DIFFUSE_SPECULAR_SH (NRD_SIGNAL = BOTH, NRD_MODE = SH) restated from /root/reference/External/NRD/Shaders:
//   RELAX_ClassifyTiles.cs.hlsl:21-51, RELAX_PrePass.cs.hlsl:21-385, RELAX_TemporalAccumulation.cs.hlsl:21-942,
//   RELAX_HistoryFix.cs.hlsl:21-163, RELAX_HistoryClamping.cs.hlsl:21-354, RELAX_Copy.cs.hlsl:21-34,
//   RELAX_AntiFirefly.cs.hlsl:21-216, RELAX_AtrousSmem.cs.hlsl:21-484, RELAX_Atrous.cs.hlsl:21-260,
//   helpers RELAX_Common.hlsli:11-185, constants RELAX_Config.hlsli:11-102.
// Build switches as in the reference's default build; checkerboard, confidence and disocclusion-threshold-mix inputs are restated
// (the executor rejects those), NRD_USE_PREV_WORLD_SPACE_MATRIX = 0, R10G10B10A2 normals (material IDs on).
// Shared-memory tiles of the shaders hold f( clamp( pos, 0, rectSize - 1 ) ); the restatement reads the clamped texel.
#include <cmath>
#include <string>

#include "reblur_shared.h"

namespace orc {

namespace {

struct RelaxCB {  // RELAX_Config.hlsli:21-101 (+ gStepSize, gIsLastPass of the a-trous passes), 720 bytes
    float4x4 gWorldToClip, gWorldToClipPrev, gWorldToViewPrev, gWorldPrevToWorld;
    float4 gRotatorPre, gFrustumRight, gFrustumUp, gFrustumForward, gPrevFrustumRight, gPrevFrustumUp, gPrevFrustumForward, gCameraDelta, gMvScale;
    float2 gJitter, gResolutionScale, gRectOffset, gResourceSizeInv, gResourceSize, gRectSizeInv, gRectSizePrev, gResourceSizeInvPrev;
    uint2 gPrintfAt, gRectOrigin;
    int2 gRectSize;
    float gSpecMaxAccumulatedFrameNum, gSpecMaxFastAccumulatedFrameNum, gDiffMaxAccumulatedFrameNum, gDiffMaxFastAccumulatedFrameNum, gDisocclusionThreshold,
        gDisocclusionThresholdAlternate, gCameraAttachedReflectionMaterialID, gStrandMaterialID, gStrandThickness, gRoughnessFraction, gSpecVarianceBoost, gSplitScreen,
        gDiffBlurRadius, gSpecBlurRadius, gDepthThreshold, gLobeAngleFraction, gSpecLobeAngleSlack, gHistoryFixEdgeStoppingNormalPower, gRoughnessEdgeStoppingRelaxation,
        gNormalEdgeStoppingRelaxation, gFastHistoryClampingSigmaScale, gHistoryAccelerationAmount, gHistoryResetTemporalSigmaScale, gHistoryResetSpatialSigmaScale,
        gHistoryResetAmount, gDenoisingRange, gSpecPhiLuminance, gDiffPhiLuminance, gDiffMaxLuminanceRelativeDifference, gSpecMaxLuminanceRelativeDifference,
        gLuminanceEdgeStoppingRelaxation, gConfidenceDrivenRelaxationMultiplier, gConfidenceDrivenLuminanceEdgeStoppingRelaxation,
        gConfidenceDrivenNormalEdgeStoppingRelaxation, gDebug, gOrthoMode, gUnproject, gFramerateScale, gCheckerboardResolveAccumSpeed, gHistoryFixFrameNum,
        gHistoryFixBasePixelStride, gHistoryFixAlternatePixelStride, gHistoryFixAlternatePixelStrideMaterialID, gHistoryThreshold, gViewZScale, gMinHitDistanceWeight,
        gDiffMinMaterial, gSpecMinMaterial;
    uint32_t gRoughnessEdgeStoppingEnabled, gFrameIndex, gDiffCheckerboard, gSpecCheckerboard, gHasHistoryConfidence, gHasDisocclusionThresholdMix, gResetHistory;
    uint32_t gStepSize, gIsLastPass;
    uint32_t _pad[1];
};
static_assert(sizeof(RelaxCB) == 720, "RELAX cbuffer is 720 bytes");

const float RELAX_NORMAL_ULP = 1.5f / 255.0f;
const float RELAX_MAX_ACCUM_FRAME_NUM = 255.0f;
const float RELAX_ANTILAG_ACCELERATION_AMOUNT_SCALE = 10.0f;
const float NRD_CURVATURE_HIGH_PARALLAX_DISOCCLUSION_THRESHOLD = 0.04f;
const float NRD_MAX_ALLOWED_VIRTUAL_MOTION_ACCELERATION = 5.0f;
const float NRD_STRAND_RELAXED_DISOCCLUSION_THRESHOLD = 0.25f;

const float3 g_Poisson8[8] = {float3(-0.4706069f, -0.4427112f, +0.6461146f), float3(-0.9057375f, +0.3003471f, +0.9542373f), float3(-0.3487388f, +0.4037880f, +0.5335386f),
                              float3(+0.1023042f, +0.6439373f, +0.6520134f), float3(+0.5699277f, +0.3513750f, +0.6695386f), float3(+0.2939128f, -0.1131226f, +0.3149309f),
                              float3(+0.7836658f, -0.4208784f, +0.8895339f), float3(+0.1564120f, -0.8198990f, +0.8346850f)};  // Poisson.hlsli:40-50

inline float Luminance(float3 x) { return dot(x, float3(0.2126f, 0.7152f, 0.0722f)); }  // ml:712
inline float3 RgbToYCoCg(float3 x) { return float3(dot(x, float3(0.25f, 0.5f, 0.25f)), dot(x, float3(0.5f, 0.0f, -0.5f)), dot(x, float3(-0.25f, 0.5f, -0.25f))); }  // ml:838
inline float3 YCoCgToRgb(float3 x) { float t = x.x - x.z; return float3(t + x.y, x.x + x.z, t - x.y); }  // ml:847
inline float Pow5(float x) { return Math::Pow01(1.0f - x, 5.0f); }  // ml:1905
inline float Bayer4x4(uint32_t x, uint32_t y, uint32_t frameIndex) {  // ml:1578-1602
    uint32_t px = x & 3u, py = y & 3u;
    uint32_t b = ((py & 1u) << 2) | ((px & 1u) << 3) | ((py & 2u) >> 1) | (px & 2u);
    return (float((b + frameIndex) & 0xFu) + 0.5f) / 16.0f;
}
inline float3 abs3(float3 a) { return float3(std::fabs(a.x), std::fabs(a.y), std::fabs(a.z)); }
inline float3 sqrt3(float3 a) { return float3(std::sqrt(a.x), std::sqrt(a.y), std::sqrt(a.z)); }
inline float3 clamp3(float3 v, float3 lo, float3 hi) { return min(max(v, lo), hi); }
inline float4 clamp4(float4 v, float lo, float hi) { return min(max(v, float4(lo)), float4(hi)); }
inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// RELAX_Common.hlsli
inline float4 UnpackPrevNormalRoughness(float4 p) { return float4(_NRD_SafeNormalize(p.xyz() * 2.0f - 1.0f), p.w); }  // :11
inline float4 PackPrevNormalRoughness(float4 nr) { return float4(nr.xyz() * 0.5f + 0.5f, nr.w); }                    // :21
inline float BilinearWithCustomWeightsImmediateFloat(float s00, float s10, float s01, float s11, float4 w) {            // :30
    float o = s00 * w.x;
    o += s10 * w.y;
    o += s01 * w.z;
    o += s11 * w.w;
    float sum = dot(w, float4(1.0f));
    return sum < 0.0001f ? 0.0f : o * rcp(sum);
}
inline float3 BilinearWithCustomWeightsSH(const Tex& t, int2 p, float4 w) {  // :56
    float3 o = t.load(p.x, p.y).xyz() * w.x;
    o += t.load(p.x + 1, p.y).xyz() * w.y;
    o += t.load(p.x, p.y + 1).xyz() * w.z;
    o += t.load(p.x + 1, p.y + 1).xyz() * w.w;
    float sum = dot(w, float4(1.0f));
    return sum < 0.0001f ? float3(0.0f) : o * rcp(sum);
}
inline float GetPlaneDistanceWeight(float3 centerWorldPos, float3 centerNormal, float centerViewZ, float3 sampleWorldPos, float threshold) {  // :101
    float d = std::fabs(dot(sampleWorldPos - centerWorldPos, centerNormal));
    return d / centerViewZ > threshold ? 0.0f : 1.0f;
}
inline float GetPlaneDistanceWeight_Atrous(float3 centerWorldPos, float3 centerNormal, float3 sampleWorldPos, float threshold) {  // :108
    float d = std::fabs(dot(sampleWorldPos - centerWorldPos, centerNormal));
    return d < threshold ? 1.0f : 0.0f;
}
inline float GetSpecLobeTanHalfAngle(float roughness, float percentOfVolume = 0.75f) {  // :115
    roughness = saturate(roughness);
    percentOfVolume = saturate(percentOfVolume);
    return roughness * roughness * percentOfVolume / (1.0f - percentOfVolume + NRD_EPS);
}
inline float2 GetNormalWeightParams_ATrous(float roughness, float numFramesInHistory, float specularReprojectionConfidence, float normalEdgeStoppingRelaxation,
                                           float specularLobeAngleFraction, float specularLobeAngleSlack) {  // :125
    float relaxation = saturate(numFramesInHistory / 5.0f);
    relaxation *= lerp(1.0f, specularReprojectionConfidence, normalEdgeStoppingRelaxation);
    float f = 0.9f + 0.1f * relaxation;
    float angle = std::atan(GetSpecLobeTanHalfAngle(roughness, specularLobeAngleFraction));
    angle *= 10.0f - 9.0f * relaxation;
    angle += specularLobeAngleSlack;
    angle = min(Math::Pi(0.5f), angle);
    return float2(angle, f);
}
inline float GetSpecularNormalWeight_ATrous(float2 params0, float3 n0, float3 n, float3 v0, float3 v) {  // :147
    float cosa = min(dot(n0, n), dot(v0, v));
    float a = Math::AcosApproxPositive(cosa);
    a = Math::SmoothStep(0.0f, params0.x, a);
    return saturate(1.0f - a * params0.y);
}
inline float GetNormalWeightParam2(float roughness, float angleFraction) {  // :159
    float angle = std::atan(GetSpecLobeTanHalfAngle(roughness, angleFraction));
    return 1.0f / max(angle, RELAX_NORMAL_ULP);
}
inline float ApplyThinLensEquation(float O, float curvature) { return O / (2.0f * curvature * O + 1.0f); }  // TA:23

struct Ctx {
    const RelaxCB& cb;
    explicit Ctx(const RelaxCB& c) : cb(c) {}
    float UnpackViewZ(float z) const { return std::fabs(z * cb.gViewZScale); }
    bool IsInDenoisingRange(float z) const { return z < cb.gDenoisingRange; }
    float ApplyGeometryWeightLast(float w, float z, float NoX, float2 p) const {
        w *= ComputeWeight(NoX, p.x, p.y);
        return !IsInDenoisingRange(z) ? 0.0f : w;
    }
    float3 worldPos(float2 clipXY, float viewZ, float4 fwd, float4 right, float4 up) const {  // RELAX_Common.hlsli:69-99
        if (cb.gOrthoMode == 0.0f) return viewZ * (fwd.xyz() + right.xyz() * clipXY.x - up.xyz() * clipXY.y);
        return viewZ * fwd.xyz() + right.xyz() * clipXY.x - up.xyz() * clipXY.y;
    }
    float3 GetCurrentWorldPosFromClipSpaceXY(float2 c, float viewZ) const { return worldPos(c, viewZ, cb.gFrustumForward, cb.gFrustumRight, cb.gFrustumUp); }
    float3 GetCurrentWorldPosFromPixelPos(int px, int py, float viewZ) const {
        float2 c = (float2((float)px, (float)py) + float2(0.5f, 0.5f)) * cb.gRectSizeInv * 2.0f - 1.0f;
        return GetCurrentWorldPosFromClipSpaceXY(c, viewZ);
    }
    float3 GetPreviousWorldPosFromClipSpaceXY(float2 c, float viewZ) const { return worldPos(c, viewZ, cb.gPrevFrustumForward, cb.gPrevFrustumRight, cb.gPrevFrustumUp); }
    float3 GetPreviousWorldPosFromPixelPos(int px, int py, float viewZ) const {
        float2 c = (float2((float)px, (float)py) + float2(0.5f, 0.5f)) * (float2(1.0f) / cb.gRectSizePrev) * 2.0f - 1.0f;
        return GetPreviousWorldPosFromClipSpaceXY(c, viewZ);
    }
    float2 gRectSizeF() const { return float2((float)cb.gRectSize.x, (float)cb.gRectSize.y); }
    float2 gResolutionScalePrev() const { return cb.gRectSizePrev * cb.gResourceSizeInvPrev; }
    // ClampUvToViewport, NRD_SUPPORTS_VIEWPORT_OFFSET = 0 (Common.hlsli:242)
    float2 ClampUvToViewport(float2 uv) const { return min(uv * cb.gResolutionScale, cb.gResolutionScale - 0.5f * cb.gResourceSizeInv); }
};

float4 unpackNR(const Tex& t, int x, int y, float& materialID) { return NRD_FrontEnd_UnpackNormalAndRoughness(t.load(x, y), materialID); }
float4 unpackNR(const Tex& t, int x, int y) { float m; return unpackNR(t, x, y, m); }

// ---------------------------------------------------------------------------------------------------------------
// RELAX_ClassifyTiles.cs.hlsl:21-51
void classifyTiles(const RelaxCB& cb, const Tex& gIn_ViewZ, Tex& gOut_Tiles, int gridW, int gridH) {
    Ctx c(cb);
#pragma omp parallel for schedule(dynamic, 1)
    for (int ty = 0; ty < gridH; ty++)
        for (int tx = 0; tx < gridW; tx++) {
            int sky = 0;
            for (int j = 0; j < 16; j++)
                for (int i = 0; i < 16; i++) sky += !c.IsInDenoisingRange(std::fabs(gIn_ViewZ.load(tx * 16 + i, ty * 16 + j).x)) ? 1 : 0;
            gOut_Tiles.store(tx, ty, float4(sky == 256 ? 1.0f : 0.0f, 0, 0, 0));
        }
}

// RELAX_Common.hlsli:164-165
inline float GetBilateralWeight(float z, float zc) { return Math::LinearStep(0.03f, 0.0f, std::fabs(z - zc) * rcp(max(z, zc))); }

// RELAX_PrePass.cs.hlsl:21-385
void prePass(const RelaxCB& cb, const Tex& gIn_Tiles, const Tex& gIn_Normal_Roughness, const Tex& gIn_ViewZ, const Tex& gIn_Spec, const Tex& gIn_Diff, const Tex& gIn_SpecSh,
             const Tex& gIn_DiffSh, Tex& gOut_Spec, Tex& gOut_Diff, Tex& gOut_SpecSh, Tex& gOut_DiffSh, int gridW, int gridH) {
    Ctx c(cb);
    const float2 rectSize = c.gRectSizeF();
#pragma omp parallel for schedule(dynamic, 4)
    for (int py = 0; py < gridH * 16; py++)
        for (int px = 0; px < gridW * 16; px++) {
            if (gIn_Tiles.load(px >> 4, py >> 4).x != 0.0f || px >= cb.gRectSize.x || py >= cb.gRectSize.y) continue;
            float centerViewZ = c.UnpackViewZ(gIn_ViewZ.load(px, py).x);
            if (!c.IsInDenoisingRange(centerViewZ)) continue;

            float centerMaterialID;
            float4 centerNormalRoughness = unpackNR(gIn_Normal_Roughness, px, py, centerMaterialID);
            float3 centerNormal = centerNormalRoughness.xyz();
            float centerRoughness = centerNormalRoughness.w;
            float3 centerWorldPos = c.GetCurrentWorldPosFromPixelPos(px, py, centerViewZ);
            float4 rotator = cb.gRotatorPre;  // NRD_FRAME
            float2 pixelUv = float2(px + 0.5f, py + 0.5f) * cb.gRectSizeInv;

            // Checkerboard resolve weights ( :39-72 )
            const uint32_t checkerboard = Sequence::CheckerBoard((uint32_t)px, (uint32_t)py, cb.gFrameIndex);
            int cbX0 = std::max(px - 1, 0), cbX1 = std::min(px + 1, cb.gRectSize.x - 1);
            float materialID0 = 0.0f, materialID1 = 0.0f;
            float2 checkerboardResolveWeights = float2(1.0f);
            if (cb.gSpecCheckerboard != 2 || cb.gDiffCheckerboard != 2) {
                float viewZ0 = c.UnpackViewZ(gIn_ViewZ.load(cbX0, py).x), viewZ1 = c.UnpackViewZ(gIn_ViewZ.load(cbX1, py).x);
                unpackNR(gIn_Normal_Roughness, cbX0, py, materialID0);
                unpackNR(gIn_Normal_Roughness, cbX1, py, materialID1);
                checkerboardResolveWeights = float2(GetBilateralWeight(viewZ0, centerViewZ), GetBilateralWeight(viewZ1, centerViewZ));
                checkerboardResolveWeights.x = (!c.IsInDenoisingRange(viewZ0) || px < 1) ? 0.0f : checkerboardResolveWeights.x;
                checkerboardResolveWeights.y = (!c.IsInDenoisingRange(viewZ1) || px > cb.gRectSize.x - 2) ? 0.0f : checkerboardResolveWeights.y;
            }
            cbX0 >>= 1;
            cbX1 >>= 1;
            // ApplyCheckerboardShift ( Common.hlsli:332-342 ) on a pixel-centre position
            auto applyCheckerboardShift = [&](float2 pos, uint32_t mode, int counter) {
                float2 posPositive = pos + 16384.0f;
                uint32_t cbd = Sequence::CheckerBoard((uint32_t)posPositive.x, (uint32_t)posPositive.y, cb.gFrameIndex);
                float shift = ((counter & 0x1) == 0) ? -1.0f : 1.0f;
                pos.x += shift * float(cbd != mode && mode != 2);
                return pos;
            };

            // ---- diffuse ----
            bool diffHasData = true;
            int diffX = px;
            if (cb.gDiffCheckerboard != 2) {
                diffHasData = checkerboard == cb.gDiffCheckerboard;
                diffX >>= 1;
            }
            float4 diffuseIllumination = gIn_Diff.load(diffX, py);
            float3 diffuseSH = gIn_DiffSh.load(diffX, py).xyz();
            if (!diffHasData) {
                float2 wc = checkerboardResolveWeights;
                wc.x *= float(CompareMaterials(centerMaterialID, materialID0, cb.gDiffMinMaterial));
                wc.y *= float(CompareMaterials(centerMaterialID, materialID1, cb.gDiffMinMaterial));
                wc *= Math::PositiveRcp(wc.x + wc.y);
                float4 d0 = gIn_Diff.load(cbX0, py), d1 = gIn_Diff.load(cbX1, py);
                if (wc.x == 0.0f) d0 = float4(0.0f);
                if (wc.y == 0.0f) d1 = float4(0.0f);
                diffuseIllumination = d0 * wc.x + d1 * wc.y;
                float3 d0SH = gIn_DiffSh.load(cbX0, py).xyz(), d1SH = gIn_DiffSh.load(cbX1, py).xyz();
                if (wc.x == 0.0f) d0SH = float3(0.0f);
                if (wc.y == 0.0f) d1SH = float3(0.0f);
                diffuseSH = d0SH * wc.x + d1SH * wc.y;
            }
            if (cb.gDiffBlurRadius > 0.0f) {
                float frustumSize = PixelRadiusToWorld(cb.gUnproject, cb.gOrthoMode, (float)std::min(cb.gRectSize.x, cb.gRectSize.y), centerViewZ);
                float hitDist = diffuseIllumination.w == 0.0f ? 1.0f : diffuseIllumination.w;
                float hitDistFactor = GetHitDistFactor(hitDist, frustumSize);
                float blurRadius = cb.gDiffBlurRadius * hitDistFactor;
                if (diffuseIllumination.w == 0.0f) blurRadius = max(blurRadius, 1.0f);
                float normalWeightParam = GetNormalWeightParam2(1.0f, 0.25f * cb.gLobeAngleFraction);
                float2 hitDistanceWeightParams = GetHitDistanceWeightParams(diffuseIllumination.w, 1.0f / 9.0f);
                float weightSum = 1.0f;
                float diffMinHitDistanceWeight = cb.gMinHitDistanceWeight;
                for (int i = 0; i < 8; i++) {
                    float3 offset = g_Poisson8[i];
                    float2 uv = pixelUv * rectSize + Geometry::RotateVector(rotator, float2(offset.x, offset.y)) * blurRadius;
                    uv = floor(uv) + 0.5f;
                    uv = applyCheckerboardShift(uv, cb.gDiffCheckerboard, i);
                    uv = uv * cb.gRectSizeInv;
                    float2 uvScaled = c.ClampUvToViewport(uv);
                    float2 checkerboardUvScaled = float2(uvScaled.x * (cb.gDiffCheckerboard != 2 ? 0.5f : 1.0f), uvScaled.y);

                    float sampleMaterialID;
                    float3 sampleNormal = NRD_FrontEnd_UnpackNormalAndRoughness(gIn_Normal_Roughness.sampleNearest(uvScaled), sampleMaterialID).xyz();
                    float sampleViewZ = c.UnpackViewZ(gIn_ViewZ.sampleNearest(uvScaled).x);
                    float3 sampleWorldPos = c.GetCurrentWorldPosFromClipSpaceXY(uv * 2.0f - 1.0f, sampleViewZ);

                    float sampleWeight = IsInScreenNearest(uv);
                    sampleWeight *= float(c.IsInDenoisingRange(sampleViewZ));
                    sampleWeight *= float(CompareMaterials(centerMaterialID, sampleMaterialID, cb.gDiffMinMaterial));
                    sampleWeight *= GetPlaneDistanceWeight(centerWorldPos, centerNormal, cb.gOrthoMode == 0.0f ? centerViewZ : 1.0f, sampleWorldPos, cb.gDepthThreshold);
                    float angle = Math::AcosApproxPositive(dot(centerNormal, sampleNormal));
                    sampleWeight *= ComputeWeight(angle, normalWeightParam, 0.0f);

                    float4 sampleDiffuseIllumination = gIn_Diff.sampleNearest(checkerboardUvScaled);
                    if (sampleWeight == 0.0f) sampleDiffuseIllumination = float4(0.0f);  // Denanify
                    sampleWeight *= lerp(diffMinHitDistanceWeight, 1.0f, ComputeExponentialWeight(sampleDiffuseIllumination.w, hitDistanceWeightParams.x, hitDistanceWeightParams.y));
                    sampleWeight *= GetGaussianWeight(offset.z);

                    weightSum += sampleWeight;
                    diffuseIllumination += sampleDiffuseIllumination * sampleWeight;
                    float3 sampleDiffuseSH = gIn_DiffSh.sampleNearest(checkerboardUvScaled).xyz();
                    if (sampleWeight == 0.0f) sampleDiffuseSH = float3(0.0f);
                    diffuseSH += sampleDiffuseSH * sampleWeight;
                }
                diffuseIllumination = diffuseIllumination / weightSum;
                diffuseSH = diffuseSH / weightSum;
            }
            gOut_Diff.store(px, py, clamp4(diffuseIllumination, 0.0f, NRD_FP16_MAX));
            gOut_DiffSh.store(px, py, float4(clamp3(diffuseSH, float3(-NRD_FP16_MAX), float3(NRD_FP16_MAX)), 0.0f));

            // ---- specular ----
            RngHash rng;
            rng.Initialize((uint32_t)px, (uint32_t)py, cb.gFrameIndex);
            bool specHasData = true;
            int specX = px;
            if (cb.gSpecCheckerboard != 2) {
                specHasData = checkerboard == cb.gSpecCheckerboard;
                specX >>= 1;
            }
            float4 specularIllumination = gIn_Spec.load(specX, py);
            float3 specularSH = gIn_SpecSh.load(specX, py).xyz();
            if (!specHasData) {
                float2 wc = checkerboardResolveWeights;
                wc.x *= float(CompareMaterials(centerMaterialID, materialID0, cb.gSpecMinMaterial));
                wc.y *= float(CompareMaterials(centerMaterialID, materialID1, cb.gSpecMinMaterial));
                wc *= Math::PositiveRcp(wc.x + wc.y);
                float4 s0 = gIn_Spec.load(cbX0, py), s1 = gIn_Spec.load(cbX1, py);
                if (wc.x == 0.0f) s0 = float4(0.0f);
                if (wc.y == 0.0f) s1 = float4(0.0f);
                specularIllumination = s0 * wc.x + s1 * wc.y;
                float3 s0SH = gIn_SpecSh.load(cbX0, py).xyz(), s1SH = gIn_SpecSh.load(cbX1, py).xyz();
                if (wc.x == 0.0f) s0SH = float3(0.0f);
                if (wc.y == 0.0f) s1SH = float3(0.0f);
                specularSH = s0SH * wc.x + s1SH * wc.y;
            }
            specularIllumination.w = max(0.0f, min(cb.gDenoisingRange, specularIllumination.w));
            if (cb.gSpecBlurRadius > 0.0f) {
                float3 viewVector = cb.gOrthoMode == 0.0f ? normalize(-centerWorldPos) : cb.gFrustumForward.xyz();
                float4 D = ImportanceSampling::GetSpecularDominantDirectionG2(centerNormal, viewVector, centerRoughness);
                float NoD = std::fabs(dot(centerNormal, D.xyz()));
                float frustumSize = PixelRadiusToWorld(cb.gUnproject, cb.gOrthoMode, (float)std::min(cb.gRectSize.x, cb.gRectSize.y), centerViewZ);
                float hitDist = specularIllumination.w == 0.0f ? 1.0f : specularIllumination.w;
                float hitDistFactor = GetHitDistFactor(hitDist * NoD, frustumSize);
                float smc = GetSpecMagicCurve(centerRoughness);
                float blurRadius = cb.gSpecBlurRadius * hitDistFactor * smc;
                float lobeTanHalfAngle = ImportanceSampling::GetSpecularLobeTanHalfAngle(centerRoughness);
                float lobeRadius = hitDist * NoD * lobeTanHalfAngle;
                float minBlurRadius = lobeRadius / PixelRadiusToWorld(cb.gUnproject, cb.gOrthoMode, 1.0f, centerViewZ + hitDist * D.w);
                blurRadius = min(blurRadius, minBlurRadius);
                if (specularIllumination.w == 0.0f) blurRadius = max(blurRadius, 1.0f);

                float normalWeightParam = GetNormalWeightParam2(centerRoughness, 0.5f * cb.gLobeAngleFraction);
                float2 hitDistanceWeightParams = GetHitDistanceWeightParams(specularIllumination.w, 1.0f / 9.0f);
                float2 roughnessWeightParams = GetRoughnessWeightParams(centerRoughness, cb.gRoughnessFraction);
                float specMinHitDistanceWeight = specularIllumination.w == 0.0f ? 1.0f : cb.gMinHitDistanceWeight * smc;
                float specularHitT = specularIllumination.w == 0.0f ? cb.gDenoisingRange : specularIllumination.w;
                float NoV = std::fabs(dot(centerNormal, viewVector));
                float minHitT = specularHitT == 0.0f ? NRD_INF : specularHitT;
                float weightSum = 1.0f;
                for (int i = 0; i < 8; i++) {
                    float3 offset = g_Poisson8[i];
                    float2 uv = pixelUv * rectSize + Geometry::RotateVector(rotator, float2(offset.x, offset.y)) * blurRadius;
                    uv = floor(uv) + 0.5f;
                    uv = applyCheckerboardShift(uv, cb.gSpecCheckerboard, i);
                    uv = uv * cb.gRectSizeInv;
                    float2 uvScaled = c.ClampUvToViewport(uv);
                    float2 checkerboardUvScaled = float2(uvScaled.x * (cb.gSpecCheckerboard != 2 ? 0.5f : 1.0f), uvScaled.y);

                    float sampleMaterialID;
                    float4 sampleNormalRoughness = NRD_FrontEnd_UnpackNormalAndRoughness(gIn_Normal_Roughness.sampleNearest(uvScaled), sampleMaterialID);
                    float3 sampleNormal = sampleNormalRoughness.xyz();
                    float sampleRoughness = sampleNormalRoughness.w;
                    float sampleViewZ = c.UnpackViewZ(gIn_ViewZ.sampleNearest(uvScaled).x);

                    float sampleWeight = IsInScreenNearest(uv);
                    sampleWeight *= float(c.IsInDenoisingRange(sampleViewZ));
                    sampleWeight *= float(CompareMaterials(centerMaterialID, sampleMaterialID, cb.gSpecMinMaterial));
                    sampleWeight *= ComputeWeight(sampleRoughness, roughnessWeightParams.x, roughnessWeightParams.y);
                    float angle = Math::AcosApproxPositive(dot(centerNormal, sampleNormal));
                    sampleWeight *= ComputeWeight(angle, normalWeightParam, 0.0f);
                    float3 sampleWorldPos = c.GetCurrentWorldPosFromClipSpaceXY(uv * 2.0f - 1.0f, sampleViewZ);
                    sampleWeight *= GetPlaneDistanceWeight(centerWorldPos, centerNormal, cb.gOrthoMode == 0.0f ? centerViewZ : 1.0f, sampleWorldPos, cb.gDepthThreshold);

                    float4 sampleSpecularIllumination = gIn_Spec.sampleNearest(checkerboardUvScaled);
                    if (sampleWeight == 0.0f) sampleSpecularIllumination = float4(0.0f);
                    if (rng.GetFloat() < sampleWeight * NoV) minHitT = min(minHitT, sampleSpecularIllumination.w == 0.0f ? NRD_INF : sampleSpecularIllumination.w);

                    sampleWeight *= lerp(specMinHitDistanceWeight, 1.0f, ComputeExponentialWeight(sampleSpecularIllumination.w, hitDistanceWeightParams.x, hitDistanceWeightParams.y));
                    sampleWeight *= GetGaussianWeight(offset.z);

                    // less weight for samples that most likely sit at the reflection contact
                    float d = length(sampleWorldPos - centerWorldPos);
                    float h = sampleSpecularIllumination.w;
                    float t = h / (specularIllumination.w + d);
                    sampleWeight *= lerp(saturate(t), 1.0f, Math::LinearStep(0.5f, 1.0f, centerRoughness));

                    weightSum += sampleWeight;
                    float3 rgb = specularIllumination.xyz() + sampleSpecularIllumination.xyz() * sampleWeight;
                    specularIllumination = float4(rgb, specularIllumination.w);
                    float3 sampleSpecularSH = gIn_SpecSh.sampleNearest(checkerboardUvScaled).xyz();
                    if (sampleWeight == 0.0f) sampleSpecularSH = float3(0.0f);
                    specularSH += sampleSpecularSH * sampleWeight;
                }
                specularIllumination = float4(specularIllumination.xyz() / weightSum, minHitT == NRD_INF ? 0.0f : minHitT);
                specularSH = specularSH / weightSum;
            }
            gOut_Spec.store(px, py, clamp4(specularIllumination, 0.0f, NRD_FP16_MAX));
            gOut_SpecSh.store(px, py, float4(clamp3(specularSH, float3(-NRD_FP16_MAX), float3(NRD_FP16_MAX)), 0.0f));
        }
}

// ---------------------------------------------------------------------------------------------------------------
// RELAX_TemporalAccumulation.cs.hlsl
struct TaTex {
    const Tex *gIn_Tiles, *gIn_Mv, *gIn_Normal_Roughness, *gIn_ViewZ, *gIn_DisocclusionThresholdMix, *gPrev_Normal_Roughness, *gPrev_ViewZ, *gPrev_HistoryLength,
        *gPrev_MaterialID, *gIn_Spec, *gIn_Diff, *gHistory_SpecFast, *gHistory_DiffFast, *gHistory_Spec, *gHistory_Diff, *gPrev_SpecHitDist, *gIn_SpecConfidence,
        *gIn_DiffConfidence, *gIn_SpecSh, *gIn_DiffSh, *gHistory_SpecShFast, *gHistory_DiffShFast, *gHistory_SpecSh, *gHistory_DiffSh;
    Tex *gOut_HistoryLength, *gOut_Spec, *gOut_Diff, *gOut_SpecFast, *gOut_DiffFast, *gOut_SpecHitDist, *gOut_SpecReprojectionConfidence, *gOut_SpecSh, *gOut_DiffSh,
        *gOut_SpecShFast, *gOut_DiffShFast;
};

// 2x2 footprint whose top-left texel is (x0, y0), clamp addressing, order 00 10 01 11 (== GatherRed( uv ).wzxy with uv = ( origin + 1 ) / size)
float4 gather4(const Tex& t, int x0, int y0) {
    return float4(t.fetchClamped(x0, y0).x, t.fetchClamped(x0 + 1, y0).x, t.fetchClamped(x0, y0 + 1).x, t.fetchClamped(x0 + 1, y0 + 1).x);
}

void temporalAccumulation(const RelaxCB& cb, const TaTex& t, int gridW, int gridH) {
    Ctx c(cb);
    const float2 rectSize = c.gRectSizeF();
    const float minRectDim = (float)std::min(cb.gRectSize.x, cb.gRectSize.y);
    // Preload( ): { normal, roughness -> spec hitT } at the rect-clamped position
    auto preload = [&](int x, int y) {
        int gx = clampi(x, 0, cb.gRectSize.x - 1), gy = clampi(y, 0, cb.gRectSize.y - 1);
        float4 nr = NRD_FrontEnd_UnpackNormalAndRoughness(t.gIn_Normal_Roughness->load(gx, gy));
        return float4(nr.xyz(), t.gIn_Spec->load(gx, gy).w);
    };
#pragma omp parallel for schedule(dynamic, 4)
    for (int py = 0; py < gridH * 16; py++)
        for (int px = 0; px < gridW * 8; px++) {
            if (t.gIn_Tiles->load(px >> 4, py >> 4).x != 0.0f || px >= cb.gRectSize.x || py >= cb.gRectSize.y) continue;
            float currentLinearZ = c.UnpackViewZ(t.gIn_ViewZ->load(px, py).x);
            if (!c.IsInDenoisingRange(currentLinearZ)) continue;

            float currentMaterialID;
            float4 currentNormalRoughness = unpackNR(*t.gIn_Normal_Roughness, px, py, currentMaterialID);
            float3 currentNormal = currentNormalRoughness.xyz();
            float currentRoughness = currentNormalRoughness.w;

            float3 currentWorldPos = c.GetCurrentWorldPosFromPixelPos(px, py, currentLinearZ);
            float3 currentViewVector = cb.gOrthoMode == 0.0f ? currentWorldPos : currentLinearZ * normalize(cb.gFrustumForward.xyz());
            float3 V = -normalize(currentViewVector);
            float NoV = std::fabs(dot(currentNormal, V));

            float2 pixelUv = float2(px + 0.5f, py + 0.5f) * cb.gRectSizeInv;
            float4 mvRaw = t.gIn_Mv->load(px, py);
            float3 mv = float3(mvRaw.x, mvRaw.y, mvRaw.z) * cb.gMvScale.xyz();
            float3 prevWorldPos = currentWorldPos;
            float2 prevUVSMB = pixelUv + float2(mv.x, mv.y);
            if (cb.gMvScale.w == 0.0f) {
                if (cb.gMvScale.z == 0.0f) mv.z = Geometry::AffineTransform(cb.gWorldToViewPrev, currentWorldPos).z - currentLinearZ;
                prevWorldPos = c.GetPreviousWorldPosFromClipSpaceXY(prevUVSMB * 2.0f - 1.0f, currentLinearZ + mv.z) + cb.gCameraDelta.xyz();
            } else {
                prevWorldPos = prevWorldPos + mv;
                prevUVSMB = Geometry::GetScreenUv(cb.gWorldToClipPrev, prevWorldPos);
            }

            float3 diffuseIllumination = t.gIn_Diff->load(px, py).xyz();
            float3 diffuseSH = t.gIn_DiffSh->load(px, py).xyz();
            float4 specularIllumination = t.gIn_Spec->load(px, py);
            float3 specularSH = t.gIn_SpecSh->load(px, py).xyz();

            // average normal, min hit distance in 3x3
            float hitTM1 = preload(px, py).w;
            float minHitDist3x3 = hitTM1 == 0.0f ? NRD_INF : hitTM1;
            float3 currentNormalAveraged = currentNormal;
            for (int i = -1; i <= 1; i++)
                for (int j = -1; j <= 1; j++) {
                    if (i == 0 && j == 0) continue;
                    float4 n = preload(px + i, py + j);
                    minHitDist3x3 = min(minHitDist3x3, n.w == 0.0f ? NRD_INF : n.w);
                    currentNormalAveraged += n.xyz();
                }
            currentNormalAveraged = currentNormalAveraged / 9.0f;
            float currentRoughnessModified = Filtering::GetModifiedRoughnessFromNormalVariance(currentRoughness, currentNormalAveraged);

            float specular1stMoment = Luminance(specularIllumination.xyz());
            float specular2ndMoment = specular1stMoment * specular1stMoment;
            float diffuse1stMoment = Luminance(diffuseIllumination);
            float diffuse2ndMoment = diffuse1stMoment * diffuse1stMoment;

            float smbParallaxInPixels1 = ComputeParallaxInPixels(prevWorldPos + cb.gCameraDelta.xyz(), cb.gOrthoMode == 0.0f ? prevUVSMB : pixelUv, cb.gWorldToClipPrev, rectSize);
            float smbParallaxInPixels2 = ComputeParallaxInPixels(prevWorldPos - cb.gCameraDelta.xyz(), cb.gOrthoMode == 0.0f ? pixelUv : prevUVSMB, cb.gWorldToClip, rectSize);
            float smbParallaxInPixelsMax = max(smbParallaxInPixels1, smbParallaxInPixels2);
            float smbParallaxInPixelsMin = min(smbParallaxInPixels1, smbParallaxInPixels2);
            float pixelSize = PixelRadiusToWorld(cb.gUnproject, cb.gOrthoMode, 1.0f, currentLinearZ);

            float disocclusionThresholdMix = 0.0f;
            if (currentMaterialID == cb.gStrandMaterialID) disocclusionThresholdMix = NRD_GetNormalizedStrandThickness(cb.gStrandThickness, pixelSize);
            if (cb.gHasDisocclusionThresholdMix) disocclusionThresholdMix = t.gIn_DisocclusionThresholdMix->load(px, py).x;
            float disocclusionThreshold = lerp(cb.gDisocclusionThreshold, cb.gDisocclusionThresholdAlternate, disocclusionThresholdMix);
            if (currentMaterialID == cb.gStrandMaterialID) {
                float mediumParallax = Math::SmoothStep01(smbParallaxInPixelsMax);
                disocclusionThreshold = lerp(NRD_STRAND_RELAXED_DISOCCLUSION_THRESHOLD, disocclusionThreshold, mediumParallax);
            }

            // ---- loadSurfaceMotionBasedPrevData (TA:48-236) ----
            float footprintQuality, historyLength, SMBReprojectionFound;
            float4 prevDiffuseIlluminationAnd2ndMomentSMB, prevSpecularIlluminationAnd2ndMomentSMB;
            float3 prevDiffuseIlluminationAnd2ndMomentSMBResponsive, prevSpecularIlluminationAnd2ndMomentSMBResponsive;
            float3 prevDiffuseSH, prevDiffuseResponsiveSH, prevSpecularSMBSH, prevSpecularSMBResponsiveSH;
            float prevReflectionHitTSMB;
            {
                float3 smbNormal = normalize(currentNormalAveraged);
                float2 prevPixelPosFloat = prevUVSMB * cb.gRectSizePrev;
                float2 fl = floor(prevPixelPosFloat - 0.5f);
                int2 bilinearOrigin = int2((int)fl.x, (int)fl.y);
                float2 bilinearWeights = frac(prevPixelPosFloat - 0.5f);

                int ox = bilinearOrigin.x, oy = bilinearOrigin.y;
                // gather uv = ( origin + { 0, 2 } ) / size -> footprints with top-left texels origin + { -1, +1 }
                float4 z00 = gather4(*t.gPrev_ViewZ, ox - 1, oy - 1), z10 = gather4(*t.gPrev_ViewZ, ox + 1, oy - 1), z01 = gather4(*t.gPrev_ViewZ, ox - 1, oy + 1),
                       z11 = gather4(*t.gPrev_ViewZ, ox + 1, oy + 1);
                auto unpack4 = [&](float4 z) { return float4(c.UnpackViewZ(z.x), c.UnpackViewZ(z.y), c.UnpackViewZ(z.z), c.UnpackViewZ(z.w)); };
                float4 prevViewZs00 = unpack4(z00), prevViewZs10 = unpack4(z10), prevViewZs01 = unpack4(z01), prevViewZs11 = unpack4(z11);
                float4 m00 = gather4(*t.gPrev_MaterialID, ox - 1, oy - 1) * 255.0f, m10 = gather4(*t.gPrev_MaterialID, ox + 1, oy - 1) * 255.0f,
                       m01 = gather4(*t.gPrev_MaterialID, ox - 1, oy + 1) * 255.0f, m11 = gather4(*t.gPrev_MaterialID, ox + 1, oy + 1) * 255.0f;

                float frustumSize = pixelSize * minRectDim;
                float slopeScale = 1.0f / lerp(lerp(0.05f, 1.0f, NoV), 1.0f, saturate(smbParallaxInPixelsMax / 30.0f));
                float4 smbDisocclusionThreshold = float4(saturate(disocclusionThreshold * slopeScale) * frustumSize);
                smbDisocclusionThreshold = smbDisocclusionThreshold * IsInScreenBilinear(float2((float)ox, (float)oy), cb.gRectSizePrev);
                smbDisocclusionThreshold -= NRD_EPS;

                float3 prevViewPos = Geometry::AffineTransform(cb.gWorldToViewPrev, prevWorldPos);
                float pz = prevViewPos.z;
                auto valid = [&](float z, float thr, float mat) {
                    float v = step(std::fabs(z - pz), thr);
                    return v * float(CompareMaterials(currentMaterialID, mat, min(cb.gSpecMinMaterial, cb.gDiffMinMaterial)));
                };
                // .yzw of 00, .xzw of 10, .xyw of 01, .xyz of 11 (the 12 taps of the bicubic footprint without corners)
                float3 tapsValid0 = float3(valid(prevViewZs00.y, smbDisocclusionThreshold.x, m00.y), valid(prevViewZs00.z, smbDisocclusionThreshold.x, m00.z),
                                           valid(prevViewZs00.w, smbDisocclusionThreshold.x, m00.w));
                float3 tapsValid1 = float3(valid(prevViewZs10.x, smbDisocclusionThreshold.y, m10.x), valid(prevViewZs10.z, smbDisocclusionThreshold.y, m10.z),
                                           valid(prevViewZs10.w, smbDisocclusionThreshold.y, m10.w));
                float3 tapsValid2 = float3(valid(prevViewZs01.x, smbDisocclusionThreshold.z, m01.x), valid(prevViewZs01.y, smbDisocclusionThreshold.z, m01.y),
                                           valid(prevViewZs01.w, smbDisocclusionThreshold.z, m01.w));
                float3 tapsValid3 = float3(valid(prevViewZs11.x, smbDisocclusionThreshold.w, m11.x), valid(prevViewZs11.y, smbDisocclusionThreshold.w, m11.y),
                                           valid(prevViewZs11.z, smbDisocclusionThreshold.w, m11.z));
                float bicubicFootprintValid = dot(tapsValid0 + tapsValid1 + tapsValid2 + tapsValid3, float3(1.0f)) > 11.5f ? 1.0f : 0.0f;
                float4 bilinearTapsValid = float4(tapsValid0.z, tapsValid1.y, tapsValid2.y, tapsValid3.x);

                float2 uv = (float2((float)ox, (float)oy) + float2(1.0f, 1.0f)) * cb.gResourceSizeInvPrev;
                float3 prevNormalFlat = UnpackPrevNormalRoughness(t.gPrev_Normal_Roughness->sampleLinear(uv)).xyz();
                if (dot(smbNormal, prevNormalFlat) < 0.0f) {
                    bilinearTapsValid = float4(0.0f);
                    bicubicFootprintValid = 0.0f;
                }

                Filtering::Bilinear bilinear;
                bilinear.weights = bilinearWeights;
                float4 bilinearCustomWeights = Filtering::GetBilinearCustomWeights(bilinear, bilinearTapsValid);
                bool useBicubic = bicubicFootprintValid > 0.0f;

                HistoryFilter hf(prevPixelPosFloat, cb.gResourceSizeInvPrev, bilinearCustomWeights, useBicubic);
                prevDiffuseIlluminationAnd2ndMomentSMB = max(hf.color(*t.gHistory_Diff), float4(0.0f));
                prevSpecularIlluminationAnd2ndMomentSMB = max(hf.color(*t.gHistory_Spec), float4(0.0f));
                prevDiffuseIlluminationAnd2ndMomentSMBResponsive = max(hf.color(*t.gHistory_DiffFast).xyz(), float3(0.0f));
                prevSpecularIlluminationAnd2ndMomentSMBResponsive = max(hf.color(*t.gHistory_SpecFast).xyz(), float3(0.0f));

                prevDiffuseSH = BilinearWithCustomWeightsSH(*t.gHistory_DiffSh, bilinearOrigin, bilinearCustomWeights);
                prevDiffuseResponsiveSH = BilinearWithCustomWeightsSH(*t.gHistory_DiffShFast, bilinearOrigin, bilinearCustomWeights);
                prevSpecularSMBSH = BilinearWithCustomWeightsSH(*t.gHistory_SpecSh, bilinearOrigin, bilinearCustomWeights);
                prevSpecularSMBResponsiveSH = BilinearWithCustomWeightsSH(*t.gHistory_SpecShFast, bilinearOrigin, bilinearCustomWeights);

                float4 prevHistoryLengths = gather4(*t.gPrev_HistoryLength, ox, oy);
                historyLength = 255.0f * BilinearWithCustomWeightsImmediateFloat(prevHistoryLengths.x, prevHistoryLengths.y, prevHistoryLengths.z, prevHistoryLengths.w, bilinearCustomWeights);
                float4 prevReflectionHitTs = gather4(*t.gPrev_SpecHitDist, ox, oy);
                prevReflectionHitTSMB = BilinearWithCustomWeightsImmediateFloat(prevReflectionHitTs.x, prevReflectionHitTs.y, prevReflectionHitTs.z, prevReflectionHitTs.w, bilinearCustomWeights);
                prevReflectionHitTSMB = max(0.001f, prevReflectionHitTSMB);

                SMBReprojectionFound = bicubicFootprintValid > 0.0f ? 2.0f : 1.0f;
                footprintQuality = bicubicFootprintValid > 0.0f ? 1.0f : dot(bilinearCustomWeights, float4(1.0f));
                if (!(bilinearTapsValid.x != 0.0f || bilinearTapsValid.y != 0.0f || bilinearTapsValid.z != 0.0f || bilinearTapsValid.w != 0.0f)) {
                    SMBReprojectionFound = 0.0f;
                    footprintQuality = 0.0f;
                }
            }

            historyLength = historyLength + 1.0f;
            historyLength = min(RELAX_MAX_ACCUM_FRAME_NUM, historyLength);

            float3 Vprev = cb.gOrthoMode == 0.0f ? -normalize(prevWorldPos - cb.gCameraDelta.xyz()) : -normalize(cb.gPrevFrustumForward.xyz());
            float NoVprev = std::fabs(dot(currentNormal, Vprev));
            float sizeQuality = (NoVprev + 1e-3f) / (NoV + 1e-3f);
            sizeQuality *= sizeQuality;
            sizeQuality *= sizeQuality;
            footprintQuality *= lerp(0.1f, 1.0f, saturate(sizeQuality + std::fabs(cb.gOrthoMode)));
            if (footprintQuality < 1.0f) {
                historyLength *= std::sqrt(footprintQuality);
                historyLength = max(historyLength, 1.0f);
            }
            historyLength = cb.gResetHistory != 0 ? 1.0f : historyLength;
            float maxAccumulatedFrameNum = 1.0f + max(cb.gDiffMaxAccumulatedFrameNum, cb.gSpecMaxAccumulatedFrameNum);
            historyLength = min(historyLength, maxAccumulatedFrameNum);

            const uint32_t checkerboard = Sequence::CheckerBoard((uint32_t)px, (uint32_t)py, cb.gFrameIndex);

            // ---- diffuse ----
            {
                float diffMaxAccumulatedFrameNum = cb.gDiffMaxAccumulatedFrameNum, diffMaxFastAccumulatedFrameNum = cb.gDiffMaxFastAccumulatedFrameNum;
                if (cb.gHasHistoryConfidence) {
                    float inDiffConfidence = saturate(t.gIn_DiffConfidence->sampleLinear(prevUVSMB).x);
                    diffMaxAccumulatedFrameNum *= inDiffConfidence;
                    diffMaxFastAccumulatedFrameNum *= inDiffConfidence;
                }
                float diffuseAlpha = SMBReprojectionFound > 0.0f ? max(1.0f / (diffMaxAccumulatedFrameNum + 1.0f), 1.0f / historyLength) : 1.0f;
                float diffuseAlphaResponsive = SMBReprojectionFound > 0.0f ? max(1.0f / (diffMaxFastAccumulatedFrameNum + 1.0f), 1.0f / historyLength) : 1.0f;
                const bool diffHasData = cb.gDiffCheckerboard == 2 || checkerboard == cb.gDiffCheckerboard;
                if (!diffHasData && historyLength > 1.0f) {
                    diffuseAlpha *= 1.0f - cb.gCheckerboardResolveAccumSpeed;
                    diffuseAlphaResponsive *= 1.0f - cb.gCheckerboardResolveAccumSpeed;
                }
                float4 accumulated = lerp(prevDiffuseIlluminationAnd2ndMomentSMB, float4(diffuseIllumination, diffuse2ndMoment), diffuseAlpha);
                float3 accumulatedResponsive = lerp(prevDiffuseIlluminationAnd2ndMomentSMBResponsive, diffuseIllumination, diffuseAlphaResponsive);
                t.gOut_Diff->store(px, py, accumulated);
                t.gOut_DiffFast->store(px, py, float4(accumulatedResponsive, 0.0f));
                t.gOut_DiffSh->store(px, py, float4(lerp(prevDiffuseSH, diffuseSH, diffuseAlpha), 0.0f));
                t.gOut_DiffShFast->store(px, py, float4(lerp(prevDiffuseResponsiveSH, diffuseSH, diffuseAlphaResponsive), 0.0f));
            }
            t.gOut_HistoryLength->store(px, py, float4(historyLength / 255.0f, 0, 0, 0));

            // ---- specular ----
            float specMaxAccumulatedFrameNum = cb.gSpecMaxAccumulatedFrameNum, specMaxFastAccumulatedFrameNum = cb.gSpecMaxFastAccumulatedFrameNum;
            if (cb.gHasHistoryConfidence) {
                float inSpecConfidence = saturate(t.gIn_SpecConfidence->sampleLinear(prevUVSMB).x);
                specMaxAccumulatedFrameNum *= inSpecConfidence;
                specMaxFastAccumulatedFrameNum *= inSpecConfidence;
            }
            float specHistoryFrames = min(specMaxAccumulatedFrameNum, historyLength);
            float specHistoryResponsiveFrames = min(specMaxFastAccumulatedFrameNum, historyLength);
            float hitDist = minHitDist3x3 == NRD_INF ? 0.0f : minHitDist3x3;

            // curvature along the direction of motion (TA:633-717)
            float curvature = 0.0f;
            {
                float2 uvForZeroParallax = cb.gOrthoMode == 0.0f ? prevUVSMB : pixelUv;
                float2 deltaUv = uvForZeroParallax - Geometry::GetScreenUv(cb.gWorldToClipPrev, prevWorldPos + cb.gCameraDelta.xyz());
                deltaUv *= rectSize;
                deltaUv /= max(smbParallaxInPixels1, 1.0f / 256.0f);

                auto edgePoint = [&](float2 d, float3& xOut) {
                    float3 x = c.GetCurrentWorldPosFromClipSpaceXY((pixelUv + d * cb.gRectSizeInv) * 2.0f - 1.0f, 1.0f);
                    float3 v = cb.gOrthoMode == 0.0f ? normalize(-x) : cb.gFrustumForward.xyz();
                    float3 o = cb.gOrthoMode == 0.0f ? float3(0.0f) : x;
                    xOut = o + v * dot(currentWorldPos - o, currentNormal) / dot(currentNormal, v);  // line-plane intersection
                };
                float3 x10, x01;
                edgePoint(float2(1.0f, 0.0f), x10);
                edgePoint(float2(0.0f, 1.0f), x01);
                float3 n10 = preload(px + 1, py).xyz(), n01 = preload(px, py + 1).xyz();

                float2 w = abs(deltaUv) + 1.0f / 256.0f;
                w /= w.x + w.y;
                float3 x = x10 * w.x + x01 * w.y;
                float3 n = normalize(n10 * w.x + n01 * w.y);

                float dither = Bayer4x4((uint32_t)px, (uint32_t)py, cb.gFrameIndex);
                float edgeFix = 1.0f - Pow5(NoV);
                float deltaUvLenFixed = smbParallaxInPixelsMin;
                deltaUvLenFixed *= 1.0f + edgeFix * (1.0f + cb.gFramerateScale * dither);
                float2 motionUvHigh = pixelUv + deltaUvLenFixed * deltaUv * cb.gRectSizeInv;
                motionUvHigh = (floor(motionUvHigh * rectSize) + 0.5f) * cb.gRectSizeInv;
                if (deltaUvLenFixed > 1.0f && IsInScreenNearest(motionUvHigh) != 0.0f) {
                    float2 uvScaled = c.ClampUvToViewport(motionUvHigh);
                    float zHigh = c.UnpackViewZ(t.gIn_ViewZ->sampleNearest(uvScaled).x);
                    float3 xHigh = c.GetCurrentWorldPosFromClipSpaceXY(motionUvHigh * 2.0f - 1.0f, zHigh);
                    float3 nHigh = NRD_FrontEnd_UnpackNormalAndRoughness(t.gIn_Normal_Roughness->sampleNearest(uvScaled)).xyz();
                    float frustumSize = minRectDim * pixelSize;
                    float2 geometryWeightParams = GetGeometryWeightParams(NRD_CURVATURE_HIGH_PARALLAX_DISOCCLUSION_THRESHOLD, frustumSize, currentWorldPos, currentNormal);
                    float NoX = dot(currentNormal, xHigh);
                    float wg = c.ApplyGeometryWeightLast(1.0f, zHigh, NoX, geometryWeightParams);
                    bool cmp = wg > 0.5f;
                    n = cmp ? nHigh : n;
                    x = cmp ? xHigh : x;
                }
                float3 edge = x - currentWorldPos;
                float edgeLenSq = Math::LengthSquared(edge);
                curvature = dot(n - currentNormal, edge) / edgeLenSq;
                if (curvature < 0.0f) {
                    float2 uv1 = Geometry::GetScreenUv(cb.gWorldToClipPrev, GetXvirtual(hitDist, curvature, currentWorldPos, currentWorldPos, currentNormal, V, currentRoughness));
                    float2 uv2 = Geometry::GetScreenUv(cb.gWorldToClipPrev, currentWorldPos);
                    float a = length((uv1 - uv2) * rectSize);
                    curvature *= float(a < NRD_MAX_ALLOWED_VIRTUAL_MOTION_ACCELERATION * smbParallaxInPixelsMax + cb.gRectSizeInv.x);
                }
            }

            float3 virtualWorldPos = GetXvirtual(hitDist, curvature, currentWorldPos, prevWorldPos, currentNormal, V, currentRoughness);

            // ---- loadVirtualMotionBasedPrevData (TA:239-357) ----
            float4 prevSpecularVMB = float4(0.0f), prevSpecularVMBResponsive = float4(0.0f);
            float3 prevNormalVMB = currentNormal, prevSpecularVMBSH = float3(0.0f), prevSpecularVMBResponsiveSH = float3(0.0f);
            float prevRoughnessVMB = 0.0f, prevReflectionHitTVMB = cb.gDenoisingRange, VMBReprojectionFound;
            float2 prevUVVMB;
            {
                float4 clip = mul(cb.gWorldToClipPrev, float4(virtualWorldPos, 1.0f));
                prevUVVMB = float2(clip.x / clip.w, clip.y / clip.w) * float2(0.5f, -0.5f) + float2(0.5f, 0.5f);
                prevUVVMB = currentMaterialID == cb.gCameraAttachedReflectionMaterialID ? prevUVSMB : prevUVVMB;
                float2 prevVirtualPixelPosFloat = prevUVVMB * cb.gRectSizePrev;
                float2 fl = floor(prevVirtualPixelPosFloat - 0.5f);
                int2 bilinearOrigin = int2((int)fl.x, (int)fl.y);
                float2 bilinearWeights = frac(prevVirtualPixelPosFloat - 0.5f);
                int ox = bilinearOrigin.x, oy = bilinearOrigin.y;

                float3 cwp = currentWorldPos - cb.gCameraDelta.xyz();
                float4 vmbDisocclusionThreshold = float4(disocclusionThreshold * (cb.gOrthoMode == 0.0f ? currentLinearZ : 1.0f));
                vmbDisocclusionThreshold = vmbDisocclusionThreshold * IsInScreenBilinear(float2((float)ox, (float)oy), cb.gRectSizePrev);
                vmbDisocclusionThreshold -= NRD_EPS;

                float4 zr = gather4(*t.gPrev_ViewZ, ox, oy);
                float4 prevViewZs = float4(c.UnpackViewZ(zr.x), c.UnpackViewZ(zr.y), c.UnpackViewZ(zr.z), c.UnpackViewZ(zr.w));
                float4 prevMaterialIDs = gather4(*t.gPrev_MaterialID, ox, oy) * 255.0f;
                auto tapValid = [&](int dx, int dy, float z, float thr, float mat) {
                    float3 p = c.GetPreviousWorldPosFromPixelPos(ox + dx, oy + dy, z);
                    float v = std::fabs(dot(cwp - p, currentNormal)) > thr ? 0.0f : 1.0f;  // isReprojectionTapValid
                    return v * float(CompareMaterials(currentMaterialID, mat, cb.gSpecMinMaterial));
                };
                float4 bilinearTapsValid = float4(tapValid(0, 0, prevViewZs.x, vmbDisocclusionThreshold.x, prevMaterialIDs.x), tapValid(1, 0, prevViewZs.y, vmbDisocclusionThreshold.y, prevMaterialIDs.y),
                                                  tapValid(0, 1, prevViewZs.z, vmbDisocclusionThreshold.z, prevMaterialIDs.z), tapValid(1, 1, prevViewZs.w, vmbDisocclusionThreshold.w, prevMaterialIDs.w));
                bool anyValid = bilinearTapsValid.x != 0.0f || bilinearTapsValid.y != 0.0f || bilinearTapsValid.z != 0.0f || bilinearTapsValid.w != 0.0f;
                bool allValid = bilinearTapsValid.x != 0.0f && bilinearTapsValid.y != 0.0f && bilinearTapsValid.z != 0.0f && bilinearTapsValid.w != 0.0f;
                if (anyValid) {
                    Filtering::Bilinear bilinear;
                    bilinear.weights = bilinearWeights;
                    float4 bilinearCustomWeights = Filtering::GetBilinearCustomWeights(bilinear, bilinearTapsValid);
                    bool useBicubic = SMBReprojectionFound == 2.0f && allValid;
                    HistoryFilter hf(prevVirtualPixelPosFloat, cb.gResourceSizeInvPrev, bilinearCustomWeights, useBicubic);
                    prevSpecularVMB = max(hf.color(*t.gHistory_Spec), float4(0.0f));
                    prevSpecularVMBResponsive = max(hf.color(*t.gHistory_SpecFast), float4(0.0f));
                    prevSpecularVMBSH = BilinearWithCustomWeightsSH(*t.gHistory_SpecSh, bilinearOrigin, bilinearCustomWeights);
                    prevSpecularVMBResponsiveSH = BilinearWithCustomWeightsSH(*t.gHistory_SpecShFast, bilinearOrigin, bilinearCustomWeights);
                    prevReflectionHitTVMB = t.gPrev_SpecHitDist->sampleLinear(prevUVVMB * c.gResolutionScalePrev()).x;
                    prevReflectionHitTVMB = max(0.001f, prevReflectionHitTVMB);
                    float4 prevNormalRoughness = UnpackPrevNormalRoughness(t.gPrev_Normal_Roughness->sampleLinear(prevUVVMB * c.gResolutionScalePrev()));
                    prevNormalVMB = prevNormalRoughness.xyz();
                    prevRoughnessVMB = prevNormalRoughness.w;
                }
                VMBReprojectionFound = allValid ? 1.0f : 0.0f;
            }

            float4 D = ImportanceSampling::GetSpecularDominantDirectionG2(currentNormal, V, currentRoughnessModified);
            float virtualHistoryAmount = VMBReprojectionFound * D.w;
            virtualHistoryAmount *= cb.gOrthoMode == 0.0f ? 1.0f : 0.75f;
            virtualHistoryAmount *= float(dot(prevNormalVMB, currentNormalAveraged) > 0.0f);

            float2 uvDiff = prevUVVMB - prevUVSMB;
            float uvDiffLengthInPixels = length(uvDiff * rectSize);
            float tanCurvature = std::fabs(curvature * pixelSize);
            tanCurvature *= max(uvDiffLengthInPixels / max(NoV, 0.01f), 1.0f);
            float curvatureAngle = std::atan(tanCurvature);

            float lobeHalfAngle = max(std::atan(GetSpecLobeTanHalfAngle(currentRoughnessModified)), RELAX_NORMAL_ULP);
            float normalWeight = GetEncodingAwareNormalWeight(currentNormal, prevNormalVMB, lobeHalfAngle, curvatureAngle, RELAX_NORMAL_ULP);
            virtualHistoryAmount *= lerp(1.0f - saturate(uvDiffLengthInPixels), 1.0f, normalWeight);

            float2 relaxedRoughnessWeightParams = GetRelaxedRoughnessWeightParams(currentRoughness * currentRoughness, cb.gRoughnessFraction);
            float virtualRoughnessWeight = ComputeWeight(prevRoughnessVMB * prevRoughnessVMB, relaxedRoughnessWeightParams.x, relaxedRoughnessWeightParams.y);
            virtualRoughnessWeight = lerp(1.0f - saturate(uvDiffLengthInPixels), 1.0f, virtualRoughnessWeight);
            virtualHistoryAmount *= cb.gOrthoMode == 0.0f ? virtualRoughnessWeight : 1.0f;
            float specVMBConfidence = virtualRoughnessWeight * 0.9f + 0.1f;

            // "looking back" 1 and 2 frames
            uvDiff = uvDiff * Math::Rsqrt(Math::LengthSquared(uvDiff));
            uvDiff = uvDiff / cb.gRectSizePrev;
            uvDiff = uvDiff * (saturate(uvDiffLengthInPixels / 0.1f) + uvDiffLengthInPixels / 2.0f);
            float2 backUV1 = prevUVVMB + 1.0f * uvDiff;
            float2 backUV2 = prevUVVMB + 2.0f * uvDiff;
            float4 backNormalRoughness1 = UnpackPrevNormalRoughness(t.gPrev_Normal_Roughness->sampleLinear(backUV1 * c.gResolutionScalePrev()));
            float4 backNormalRoughness2 = UnpackPrevNormalRoughness(t.gPrev_Normal_Roughness->sampleLinear(backUV2 * c.gResolutionScalePrev()));
            float prevPrevNormalWeight = IsInScreenNearest(backUV1) != 0.0f ? GetEncodingAwareNormalWeight(prevNormalVMB, backNormalRoughness1.xyz(), lobeHalfAngle, curvatureAngle * 2.0f, RELAX_NORMAL_ULP) : 1.0f;
            prevPrevNormalWeight *= IsInScreenNearest(backUV2) != 0.0f ? GetEncodingAwareNormalWeight(prevNormalVMB, backNormalRoughness2.xyz(), lobeHalfAngle, curvatureAngle * 3.0f, RELAX_NORMAL_ULP) : 1.0f;
            virtualHistoryAmount *= 0.33f + 0.67f * prevPrevNormalWeight;
            specVMBConfidence *= 0.33f + 0.67f * prevPrevNormalWeight;
            float rw = ComputeWeight(backNormalRoughness1.w * backNormalRoughness1.w, relaxedRoughnessWeightParams.x, relaxedRoughnessWeightParams.y);
            rw *= ComputeWeight(backNormalRoughness2.w * backNormalRoughness2.w, relaxedRoughnessWeightParams.x, relaxedRoughnessWeightParams.y);
            virtualHistoryAmount *= cb.gOrthoMode == 0.0f ? rw * 0.9f + 0.1f : 1.0f;

            // virtual history confidence - hit distance
            float SMC = GetSpecMagicCurve(currentRoughnessModified);
            float hitDistC = lerp(specularIllumination.w, prevReflectionHitTSMB, SMC);
            float hitDist1 = ApplyThinLensEquation(hitDistC, curvature);
            float hitDist2 = ApplyThinLensEquation(prevReflectionHitTVMB, curvature);
            float maxDist = max(hitDist1, hitDist2);
            float dHitT = std::fabs(hitDist1 - hitDist2);
            float dHitTMultiplier = lerp(20.0f, 0.0f, SMC);
            float virtualHistoryHitDistConfidence = 1.0f - saturate(dHitTMultiplier * dHitT / (currentLinearZ + maxDist));
            virtualHistoryHitDistConfidence = lerp(virtualHistoryHitDistConfidence, 1.0f, SMC);

            // virtual history confidence - virtual UV discrepancy
            float virtualWorldPosLength = length(virtualWorldPos);
            float hitDistForTrackingPrev = prevSpecularVMBResponsive.w;
            float3 prevVirtualWorldPos = GetXvirtual(hitDistForTrackingPrev, curvature, currentWorldPos, prevWorldPos, currentNormal, V, currentRoughness);
            float virtualWorldPosLengthPrev = length(prevVirtualWorldPos);
            float2 prevUVVMBTest = Geometry::GetScreenUv(cb.gWorldToClipPrev, prevVirtualWorldPos);
            prevUVVMBTest = currentMaterialID == cb.gCameraAttachedReflectionMaterialID ? prevUVSMB : prevUVVMBTest;
            float lobeTanHalfAngle = GetSpecLobeTanHalfAngle(currentRoughness, 0.6f);
            lobeTanHalfAngle = max(lobeTanHalfAngle, 0.5f * cb.gRectSizeInv.x);
            float unproj1 = min(hitDist, hitDistForTrackingPrev) / PixelRadiusToWorld(cb.gUnproject, cb.gOrthoMode, 1.0f, max(virtualWorldPosLength, virtualWorldPosLengthPrev));
            float lobeRadiusInPixels = lobeTanHalfAngle * unproj1;
            float deltaParallaxInPixels = length((prevUVVMBTest - prevUVVMB) * rectSize);
            virtualHistoryHitDistConfidence *= Math::SmoothStep(lobeRadiusInPixels + 0.25f, 0.0f, deltaParallaxInPixels);

            // current specular signal ( surface motion )
            float specSMBConfidence = (SMBReprojectionFound > 0.0f ? 1.0f : 0.0f) * GetEncodingAwareNormalWeight(V, Vprev, lobeHalfAngle * NoV / cb.gFramerateScale, 0.0f, 0.0f);
            float specSMBAlpha = 1.0f - specSMBConfidence;
            float specSMBResponsiveAlpha = 1.0f - specSMBConfidence;
            specSMBAlpha = max(specSMBAlpha, 1.0f / (1.0f + specHistoryFrames));
            specSMBResponsiveAlpha = max(specSMBAlpha, 1.0f / (1.0f + specHistoryResponsiveFrames));
            const bool specHasData = cb.gSpecCheckerboard == 2 || checkerboard == cb.gSpecCheckerboard;
            if (!specHasData && smbParallaxInPixelsMax < 0.5f) {
                specSMBAlpha *= 1.0f - cb.gCheckerboardResolveAccumSpeed * (SMBReprojectionFound > 0.0f ? 1.0f : 0.0f);
                specSMBResponsiveAlpha *= 1.0f - cb.gCheckerboardResolveAccumSpeed * (SMBReprojectionFound > 0.0f ? 1.0f : 0.0f);
            }

            float3 accumulatedSpecularSMBrgb = lerp(prevSpecularIlluminationAnd2ndMomentSMB.xyz(), specularIllumination.xyz(), specSMBAlpha);
            float accumulatedSpecularSMBw = lerp(prevReflectionHitTSMB, specularIllumination.w, max(specSMBAlpha, 0.1f));
            float accumulatedSpecularM2SMB = lerp(prevSpecularIlluminationAnd2ndMomentSMB.w, specular2ndMoment, specSMBAlpha);
            float3 accumulatedSpecularSMBResponsive = lerp(prevSpecularIlluminationAnd2ndMomentSMBResponsive, specularIllumination.xyz(), specSMBResponsiveAlpha);

            // current specular signal ( virtual motion )
            float specVMBAlpha = 1.0f - specVMBConfidence;
            float specVMBResponsiveAlpha = 1.0f - specVMBConfidence * virtualHistoryHitDistConfidence;
            float specVMBHitTAlpha = specVMBResponsiveAlpha;
            specVMBAlpha = max(specVMBAlpha, 1.0f / (1.0f + specHistoryFrames));
            specVMBResponsiveAlpha = max(specVMBResponsiveAlpha, 1.0f / (1.0f + specHistoryResponsiveFrames));
            specVMBHitTAlpha = max(specVMBHitTAlpha, 1.0f / (1.0f + specHistoryFrames));
            if (!specHasData && smbParallaxInPixelsMax < 0.5f) {
                specVMBAlpha *= 1.0f - cb.gCheckerboardResolveAccumSpeed * (VMBReprojectionFound > 0.0f ? 1.0f : 0.0f);
                specVMBResponsiveAlpha *= 1.0f - cb.gCheckerboardResolveAccumSpeed * (VMBReprojectionFound > 0.0f ? 1.0f : 0.0f);
                specVMBHitTAlpha *= 1.0f - cb.gCheckerboardResolveAccumSpeed * (VMBReprojectionFound > 0.0f ? 1.0f : 0.0f);
            }

            float3 accumulatedSpecularVMBrgb = lerp(prevSpecularVMB.xyz(), specularIllumination.xyz(), specVMBAlpha);
            float accumulatedSpecularVMBw = lerp(prevReflectionHitTVMB, specularIllumination.w, max(specVMBHitTAlpha, 0.1f));
            float accumulatedSpecularM2VMB = lerp(prevSpecularVMB.w, specular2ndMoment, specVMBAlpha);
            float3 accumulatedSpecularVMBResponsive = lerp(prevSpecularVMBResponsive.xyz(), specularIllumination.xyz(), specVMBResponsiveAlpha);

            virtualHistoryAmount *= saturate(specVMBConfidence / (specSMBConfidence + NRD_EPS));

            float accumulatedReflectionHitT = lerp(accumulatedSpecularSMBw, accumulatedSpecularVMBw, virtualHistoryAmount);
            float3 accumulatedSpecularIllumination = lerp(accumulatedSpecularSMBrgb, accumulatedSpecularVMBrgb, virtualHistoryAmount);
            float3 accumulatedSpecularIlluminationResponsive = lerp(accumulatedSpecularSMBResponsive, accumulatedSpecularVMBResponsive, virtualHistoryAmount);
            float accumulatedSpecular2ndMoment = lerp(accumulatedSpecularM2SMB, accumulatedSpecularM2VMB, virtualHistoryAmount);

            float3 accSMBSH = lerp(prevSpecularSMBSH, specularSH, specSMBAlpha), accSMBRespSH = lerp(prevSpecularSMBResponsiveSH, specularSH, specSMBResponsiveAlpha);
            float3 accVMBSH = lerp(prevSpecularVMBSH, specularSH, specVMBAlpha), accVMBRespSH = lerp(prevSpecularVMBResponsiveSH, specularSH, specVMBResponsiveAlpha);
            t.gOut_SpecSh->store(px, py, float4(lerp(accSMBSH, accVMBSH, virtualHistoryAmount), 0.0f));
            t.gOut_SpecShFast->store(px, py, float4(lerp(accSMBRespSH, accVMBRespSH, virtualHistoryAmount), 0.0f));

            float specularHistoryConfidence = lerp(specSMBConfidence, specVMBConfidence, virtualHistoryAmount);
            if (accumulatedSpecular2ndMoment == 0.0f) accumulatedSpecular2ndMoment = cb.gSpecVarianceBoost * (1.0f - specularHistoryConfidence);

            t.gOut_Spec->store(px, py, float4(accumulatedSpecularIllumination, accumulatedSpecular2ndMoment));
            t.gOut_SpecFast->store(px, py, float4(accumulatedSpecularIlluminationResponsive, hitDist));
            t.gOut_SpecHitDist->store(px, py, float4(accumulatedReflectionHitT, 0, 0, 0));
            t.gOut_SpecReprojectionConfidence->store(px, py, float4(specularHistoryConfidence, 0, 0, 0));
        }
}

// ---------------------------------------------------------------------------------------------------------------
// RELAX_HistoryFix.cs.hlsl:21-163
void historyFix(const RelaxCB& cb, const Tex& gIn_Tiles, const Tex& gIn_HistoryLength, const Tex& gIn_Normal_Roughness, const Tex& gIn_ViewZ, const Tex& gIn_Spec,
                const Tex& gIn_Diff, const Tex& gIn_SpecSh, const Tex& gIn_DiffSh, Tex& gOut_Spec, Tex& gOut_Diff, Tex& gOut_SpecSh, Tex& gOut_DiffSh, int gridW, int gridH) {
    Ctx c(cb);
    const float2 rectSize = c.gRectSizeF();
#pragma omp parallel for schedule(dynamic, 4)
    for (int py = 0; py < gridH * 8; py++)
        for (int px = 0; px < gridW * 8; px++) {
            if (gIn_Tiles.load(px >> 4, py >> 4).x != 0.0f || px >= cb.gRectSize.x || py >= cb.gRectSize.y) continue;
            float centerViewZ = c.UnpackViewZ(gIn_ViewZ.load(px, py).x);
            float historyLength = 255.0f * gIn_HistoryLength.load(px, py).x;
            if (!c.IsInDenoisingRange(centerViewZ) || historyLength > cb.gHistoryFixFrameNum || cb.gHistoryFixFrameNum == 1.0f) continue;

            float centerMaterialID;
            float4 centerNormalRoughness = unpackNR(gIn_Normal_Roughness, px, py, centerMaterialID);
            float3 centerNormal = centerNormalRoughness.xyz();
            float centerRoughness = centerNormalRoughness.w;
            float3 centerWorldPos = c.GetCurrentWorldPosFromPixelPos(px, py, centerViewZ);
            float3 centerV = -normalize(centerWorldPos);
            float depthThreshold = cb.gDepthThreshold * (cb.gOrthoMode == 0.0f ? centerViewZ : 1.0f);

            float4 diffuseSum = gIn_Diff.load(px, py), specularSum = gIn_Spec.load(px, py);
            float3 diffuseSumSH = gIn_DiffSh.load(px, py).xyz(), specularSumSH = gIn_SpecSh.load(px, py).xyz();
            float diffuseWSum = 1.0f, specularWSum = 1.0f;
            float2 specularNormalWeightParams = GetNormalWeightParams_ATrous(centerRoughness, 5.0f, 1.0f, 0.0f, cb.gLobeAngleFraction, cb.gSpecLobeAngleSlack);

            float baseStride = centerMaterialID == cb.gHistoryFixAlternatePixelStrideMaterialID ? cb.gHistoryFixAlternatePixelStride : cb.gHistoryFixBasePixelStride;
            float r = baseStride / (1.0f + historyLength);
            r = hlsl_round(r);
            for (int j = -2; j <= 2; j++)
                for (int i = -2; i <= 2; i++) {
                    if (i == 0 && j == 0) continue;
                    // int2 * float -> float2, then truncated back to int2 by the assignment
                    int sx = (int)((float)px + (float)i * r), sy = (int)((float)py + (float)j * r);
                    float2 uv = float2(sx + 0.5f, sy + 0.5f) * cb.gRectSizeInv;
                    uv = MirrorUv(uv);
                    float2 sp = uv * rectSize;
                    sx = (int)sp.x;
                    sy = (int)sp.y;

                    float sampleMaterialID;
                    float3 sampleNormal = unpackNR(gIn_Normal_Roughness, sx, sy, sampleMaterialID).xyz();
                    float sampleViewZ = c.UnpackViewZ(gIn_ViewZ.load(sx, sy).x);
                    float3 sampleWorldPos = c.GetCurrentWorldPosFromPixelPos(sx, sy, sampleViewZ);
                    float geometryWeight = GetPlaneDistanceWeight_Atrous(centerWorldPos, centerNormal, sampleWorldPos, depthThreshold);
                    geometryWeight = c.IsInDenoisingRange(sampleViewZ) ? geometryWeight : 0.0f;

                    float diffuseW = geometryWeight;
                    diffuseW *= std::pow(max(0.01f, dot(centerNormal, sampleNormal)), max(cb.gHistoryFixEdgeStoppingNormalPower, 0.01f));
                    diffuseW *= float(CompareMaterials(sampleMaterialID, centerMaterialID, cb.gDiffMinMaterial));
                    if (diffuseW > 1e-4f) {
                        diffuseSum += gIn_Diff.load(sx, sy) * diffuseW;
                        diffuseSumSH += gIn_DiffSh.load(sx, sy).xyz() * diffuseW;
                        diffuseWSum += diffuseW;
                    }

                    float3 sampleV = -normalize(sampleWorldPos + cb.gRoughnessEdgeStoppingRelaxation * centerWorldPos);
                    float specularW = geometryWeight;
                    specularW *= GetSpecularNormalWeight_ATrous(specularNormalWeightParams, centerNormal, sampleNormal, centerV, sampleV);
                    specularW *= float(CompareMaterials(sampleMaterialID, centerMaterialID, cb.gSpecMinMaterial));
                    if (specularW > 1e-4f) {
                        specularSum += gIn_Spec.load(sx, sy) * specularW;
                        specularSumSH += gIn_SpecSh.load(sx, sy).xyz() * specularW;
                        specularWSum += specularW;
                    }
                }
            gOut_Diff.store(px, py, diffuseSum / diffuseWSum);
            gOut_DiffSh.store(px, py, float4(diffuseSumSH / diffuseWSum, 0.0f));
            gOut_Spec.store(px, py, specularSum / specularWSum);
            gOut_SpecSh.store(px, py, float4(specularSumSH / specularWSum, 0.0f));
        }
}

// RELAX_HistoryClamping.cs.hlsl:21-354
struct HcTex {
    const Tex *gIn_Tiles, *gIn_ViewZ, *gIn_HistoryLength, *gIn_SpecNoisy, *gIn_DiffNoisy, *gIn_Spec, *gIn_Diff, *gIn_SpecFast, *gIn_DiffFast, *gIn_SpecSh, *gIn_DiffSh,
        *gIn_SpecShFast, *gIn_DiffShFast;
    Tex *gOut_HistoryLength, *gOut_Spec, *gOut_Diff, *gOut_SpecFast, *gOut_DiffFast, *gOut_SpecSh, *gOut_DiffSh, *gOut_SpecShFast, *gOut_DiffShFast;
};

void historyClamping(const RelaxCB& cb, const HcTex& t, int gridW, int gridH) {
    Ctx c(cb);
    auto clampPos = [&](int& x, int& y) {
        x = clampi(x, 0, cb.gRectSize.x - 1);
        y = clampi(y, 0, cb.gRectSize.y - 1);
    };
    // one lobe of the pass; `isSpec` selects the 0.33 / 0.5 factors and what lands in the responsive alpha channel
    auto lobe = [&](int px, int py, float historyLength, bool isSpec, const Tex& noisyTex, const Tex& slowTex, const Tex& fastTex, const Tex& shTex, const Tex& shFastTex,
                    Tex& outSlow, Tex& outFast, Tex& outSh, Tex& outShFast, float maxFast, float maxSlow) {
        float3 m1 = float3(0.0f), m2 = float3(0.0f), noisyM1 = float3(0.0f);
        float noisyM2 = 0.0f, sum = 0.0f;
        for (int dx = -2; dx <= 2; dx++)
            for (int dy = -2; dy <= 2; dy++) {
                int x = px + dx, y = py + dy;
                clampPos(x, y);
                float w = float(c.IsInDenoisingRange(t.gIn_ViewZ->load(x, y).x));  // raw viewZ, as in Preload( )
                if (w != 0.0f) {
                    float3 s = RgbToYCoCg(fastTex.load(x, y).xyz());
                    m1 += s;
                    m2 += s * s;
                    float3 n = noisyTex.load(x, y).xyz();
                    float l = Luminance(n);
                    noisyM1 += n;
                    noisyM2 += l * l;
                    sum += w;
                }
            }
        m1 = m1 / sum;
        m2 = m2 / sum;
        noisyM1 = noisyM1 / sum;
        noisyM2 /= sum;
        float3 sigma = sqrt3(max(float3(0.0f), m2 - m1 * m1));
        float3 colorMin = m1 - cb.gFastHistoryClampingSigmaScale * sigma;
        float3 colorMax = m1 + cb.gFastHistoryClampingSigmaScale * sigma;

        float4 fastCenterRaw = fastTex.load(px, py);
        float4 responsiveCenterYCoCg = float4(RgbToYCoCg(fastCenterRaw.xyz()), fastCenterRaw.w);
        colorMin = min(colorMin, responsiveCenterYCoCg.xyz());
        colorMax = max(colorMax, responsiveCenterYCoCg.xyz());

        float4 slow = slowTex.load(px, py);
        float3 slowYCoCg = RgbToYCoCg(slow.xyz());
        float3 clampedYCoCg = slowYCoCg;
        if (maxFast < maxSlow) clampedYCoCg = clamp3(slowYCoCg, colorMin, colorMax);
        float3 clamped = YCoCgToRgb(clampedYCoCg);

        float4 outSlowV = float4(clamped, slow.w);
        float3 responsiveCenter = YCoCgToRgb(responsiveCenterYCoCg.xyz());
        float4 outFastV = float4(responsiveCenter, isSpec ? responsiveCenterYCoCg.w : 0.0f);
        if (historyLength <= cb.gHistoryFixFrameNum) outSlowV = isSpec ? outFastV : float4(outFastV.xyz(), outSlowV.w);

        float clampingFactor = (clampedYCoCg.x - slowYCoCg.x) == 0.0f ? 0.0f : saturate((clampedYCoCg.x - slowYCoCg.x) / (responsiveCenterYCoCg.x - slowYCoCg.x));
        if (historyLength <= cb.gHistoryFixFrameNum) clampingFactor = 1.0f;

        float historyDifferenceL = (isSpec ? 0.33f : 1.0f) * RELAX_ANTILAG_ACCELERATION_AMOUNT_SCALE * cb.gHistoryAccelerationAmount * Luminance(abs3(responsiveCenter - slow.xyz()));
        historyDifferenceL *= clampingFactor;
        if (historyLength <= cb.gHistoryFixFrameNum) historyDifferenceL = 0.0f;

        float3 distanceToNoisy = noisyM1 - responsiveCenter;
        float distanceToNoisyL = Luminance(abs3(distanceToNoisy));
        float3 acceleration = distanceToNoisyL == 0.0f ? float3(0.0f) : distanceToNoisy * historyDifferenceL / distanceToNoisyL;
        float accelerationL = Luminance(abs3(acceleration));
        float accelerationRatio = accelerationL == 0.0f ? 0.0f : distanceToNoisyL / accelerationL;
        if (accelerationRatio < 1.0f) acceleration = acceleration * accelerationRatio;
        if (accelerationRatio <= 0.0f) acceleration = float3(0.0f);
        outSlowV = float4(outSlowV.xyz() + acceleration, outSlowV.w);
        outFastV = float4(outFastV.xyz() + acceleration, outFastV.w);

        float slowL = Luminance(slow.xyz());
        float noisyInputL = Luminance(noisyM1);
        float temporalSigma = cb.gHistoryResetTemporalSigmaScale * std::sqrt(max(0.0f, noisyM2 - noisyInputL * noisyInputL));
        float spatialSigma = cb.gHistoryResetSpatialSigmaScale * sigma.x;
        float resetAmount = (isSpec ? 0.5f : 1.0f) * cb.gHistoryResetAmount * max(0.0f, std::fabs(slowL - noisyInputL) - spatialSigma - temporalSigma) /
                            (1.0e-6f + max(slowL, noisyInputL) + spatialSigma + temporalSigma);
        resetAmount = saturate(resetAmount);
        float3 noisyCenter = noisyTex.load(px, py).xyz();
        outSlowV = float4(lerp(outSlowV.xyz(), noisyCenter, resetAmount), outSlowV.w);
        outFastV = float4(lerp(outFastV.xyz(), noisyCenter, resetAmount), outFastV.w);

        float outL = Luminance(outSlowV.xyz());
        outSlowV.w += outL * outL - slowL * slowL;
        outSlowV.w = max(0.0f, outSlowV.w);

        outSlow.store(px, py, outSlowV);
        outFast.store(px, py, outFastV);
        float3 sh = shTex.load(px, py).xyz(), shFast = shFastTex.load(px, py).xyz();
        outSh.store(px, py, float4(lerp(sh, shFast, clampingFactor), 0.0f));
        outShFast.store(px, py, float4(shFast, 0.0f));
    };
#pragma omp parallel for schedule(dynamic, 4)
    for (int py = 0; py < gridH * 8; py++)
        for (int px = 0; px < gridW * 8; px++) {
            if (t.gIn_Tiles->load(px >> 4, py >> 4).x != 0.0f || px >= cb.gRectSize.x || py >= cb.gRectSize.y) continue;
            if (!c.IsInDenoisingRange(t.gIn_ViewZ->load(px, py).x)) continue;
            float historyLength = 255.0f * t.gIn_HistoryLength->load(px, py).x;
            lobe(px, py, historyLength, true, *t.gIn_SpecNoisy, *t.gIn_Spec, *t.gIn_SpecFast, *t.gIn_SpecSh, *t.gIn_SpecShFast, *t.gOut_Spec, *t.gOut_SpecFast, *t.gOut_SpecSh,
                 *t.gOut_SpecShFast, cb.gSpecMaxFastAccumulatedFrameNum, cb.gSpecMaxAccumulatedFrameNum);
            lobe(px, py, historyLength, false, *t.gIn_DiffNoisy, *t.gIn_Diff, *t.gIn_DiffFast, *t.gIn_DiffSh, *t.gIn_DiffShFast, *t.gOut_Diff, *t.gOut_DiffFast, *t.gOut_DiffSh,
                 *t.gOut_DiffShFast, cb.gDiffMaxFastAccumulatedFrameNum, cb.gDiffMaxAccumulatedFrameNum);
            t.gOut_HistoryLength->store(px, py, float4(historyLength / 255.0f, 0, 0, 0));
        }
}

// RELAX_Copy.cs.hlsl:21-34
void copy(const Tex& gIn_Spec, const Tex& gIn_Diff, Tex& gOut_Spec, Tex& gOut_Diff, int gridW, int gridH) {
#pragma omp parallel for schedule(dynamic, 8)
    for (int py = 0; py < gridH * 8; py++)
        for (int px = 0; px < gridW * 8; px++) {
            gOut_Spec.store(px, py, gIn_Spec.load(px, py));
            gOut_Diff.store(px, py, gIn_Diff.load(px, py));
        }
}

// RELAX_AntiFirefly.cs.hlsl:21-216 (cross-bilateral rank-conditioned rank-selection)
void antiFirefly(const RelaxCB& cb, const Tex& gIn_Tiles, const Tex& gIn_Normal_Roughness, const Tex& gIn_ViewZ, const Tex& gIn_Spec, const Tex& gIn_Diff, Tex& gOut_Spec,
                 Tex& gOut_Diff, int gridW, int gridH) {
    Ctx c(cb);
    auto rcrs = [&](int px, int py, const Tex& tex, float minMaterial, float centerMaterialID) {
        float4 center = tex.load(px, py);
        float centerL = Luminance(center.xyz());
        float maxL = -1.0f, minL = 1.0e6f;
        int maxX = px, maxY = py, minX = px, minY = py;
        for (int yy = -1; yy <= 1; yy++)
            for (int xx = -1; xx <= 1; xx++) {
                int x = px + xx, y = py + yy;
                if (xx == 0 && yy == 0) continue;
                if (x < 0 || y < 0 || x >= cb.gRectSize.x || y >= cb.gRectSize.y) continue;
                float sampleL = Luminance(tex.load(x, y).xyz());
                float sampleMaterialID;
                unpackNR(gIn_Normal_Roughness, x, y, sampleMaterialID);
                if (CompareMaterials(sampleMaterialID, centerMaterialID, minMaterial)) {
                    if (sampleL > maxL) { maxL = sampleL; maxX = x; maxY = y; }
                    if (sampleL < minL) { minL = sampleL; minX = x; minY = y; }
                }
            }
        int sx = px, sy = py;
        if (centerL > maxL) { sx = maxX; sy = maxY; }
        if (centerL < minL) { sx = minX; sy = minY; }
        return float4(tex.load(sx, sy).xyz(), center.w);
    };
#pragma omp parallel for schedule(dynamic, 4)
    for (int py = 0; py < gridH * 8; py++)
        for (int px = 0; px < gridW * 8; px++) {
            if (gIn_Tiles.load(px >> 4, py >> 4).x != 0.0f || px >= cb.gRectSize.x || py >= cb.gRectSize.y) continue;
            if (!c.IsInDenoisingRange(c.UnpackViewZ(gIn_ViewZ.load(px, py).x))) continue;
            float centerMaterialID;
            unpackNR(gIn_Normal_Roughness, px, py, centerMaterialID);
            gOut_Spec.store(px, py, rcrs(px, py, gIn_Spec, cb.gSpecMinMaterial, centerMaterialID));
            gOut_Diff.store(px, py, rcrs(px, py, gIn_Diff, cb.gDiffMinMaterial, centerMaterialID));
        }
}

// ---------------------------------------------------------------------------------------------------------------
// RELAX_AtrousSmem.cs.hlsl:21-484 and RELAX_Atrous.cs.hlsl:21-260
struct AtTex {
    const Tex *gIn_Tiles, *gIn_HistoryLength, *gIn_Normal_Roughness, *gIn_ViewZ, *gIn_Spec_Variance, *gIn_Diff_Variance, *gIn_SpecReprojectionConfidence, *gIn_SpecConfidence,
        *gIn_DiffConfidence, *gIn_SpecSh, *gIn_DiffSh;
    Tex *gOut_Spec_Variance, *gOut_Diff_Variance, *gOut_NormalRoughness, *gOut_MaterialID, *gOut_ViewZ, *gOut_SpecSh, *gOut_DiffSh;
};
const float kGauss3[2] = {0.44198f, 0.27901f};

void atrousSmem(const RelaxCB& cb, const AtTex& t, int gridW, int gridH) {
    Ctx c(cb);
    struct Texel { float4 spec, diff, nr; float3 specSh, diffSh, worldPos; float materialID; };
    auto fetch = [&](int x, int y) {
        int gx = clampi(x, 0, cb.gRectSize.x - 1), gy = clampi(y, 0, cb.gRectSize.y - 1);
        Texel r;
        r.spec = t.gIn_Spec_Variance->load(gx, gy);
        r.diff = t.gIn_Diff_Variance->load(gx, gy);
        r.specSh = t.gIn_SpecSh->load(gx, gy).xyz();
        r.diffSh = t.gIn_DiffSh->load(gx, gy).xyz();
        r.nr = unpackNR(*t.gIn_Normal_Roughness, gx, gy, r.materialID);
        r.worldPos = c.GetCurrentWorldPosFromPixelPos(gx, gy, c.UnpackViewZ(t.gIn_ViewZ->load(gx, gy).x));
        return r;
    };
#pragma omp parallel for schedule(dynamic, 4)
    for (int py = 0; py < gridH * 8; py++)
        for (int px = 0; px < gridW * 8; px++) {
            float isSky = t.gIn_Tiles->load(px >> 4, py >> 4).x;
            float viewZpacked = t.gIn_ViewZ->load(px, py).x;
            t.gOut_ViewZ->store(px, py, float4(viewZpacked, 0, 0, 0));
            Texel ctr = fetch(px, py);
            float4 normalRoughness = ctr.nr;
            float centerViewZ = c.UnpackViewZ(viewZpacked);
            if (!c.IsInDenoisingRange(centerViewZ)) normalRoughness = float4(1.0f / 255.0f);
            t.gOut_NormalRoughness->store(px, py, PackPrevNormalRoughness(normalRoughness));
            float3 centerWorldPos = ctr.worldPos;
            float centerMaterialID = ctr.materialID;
            t.gOut_MaterialID->store(px, py, float4(centerMaterialID / 255.0f, 0, 0, 0));

            if (isSky != 0.0f || px >= cb.gRectSize.x || py >= cb.gRectSize.y) continue;
            if (!c.IsInDenoisingRange(centerViewZ)) continue;

            float3 centerNormal = normalRoughness.xyz();
            float centerRoughness = normalRoughness.w;
            float historyLength = 255.0f * t.gIn_HistoryLength->load(px, py).x;

            if (historyLength >= cb.gHistoryThreshold) {
                // variance: 3x3 gaussian of { illumination, 2nd moment }
                const float kernel[2][2] = {{1.0f / 4.0f, 1.0f / 8.0f}, {1.0f / 8.0f, 1.0f / 16.0f}};
                float4 specularSumV = float4(0.0f), diffuseSumV = float4(0.0f);
                for (int dx = -1; dx <= 1; dx++)
                    for (int dy = -1; dy <= 1; dy++) {
                        Texel s = fetch(px + dx, py + dy);
                        float k = kernel[std::abs(dx)][std::abs(dy)];
                        specularSumV += s.spec * k;
                        diffuseSumV += s.diff * k;
                    }
                float s1 = Luminance(specularSumV.xyz()), d1 = Luminance(diffuseSumV.xyz());
                float centerSpecularVar = max(0.0f, specularSumV.w - s1 * s1), centerDiffuseVar = max(0.0f, diffuseSumV.w - d1 * d1);

                float diffuseLobeAngleFraction = cb.gLobeAngleFraction;
                float centerSpecularLuminance = Luminance(ctr.spec.xyz());
                float specularPhiLIlluminationInv = 1.0f / max(1.0e-4f, cb.gSpecPhiLuminance * std::sqrt(centerSpecularVar));
                float2 roughnessWeightParams = GetRoughnessWeightParams(centerRoughness, cb.gRoughnessFraction);
                float specularReprojectionConfidence = t.gIn_SpecReprojectionConfidence->load(px, py).x;
                float specularLuminanceWeightRelaxation = lerp(1.0f, specularReprojectionConfidence, cb.gLuminanceEdgeStoppingRelaxation);
                float diffuseLobeAngleFractionForSimplifiedSpecularNormalWeight = diffuseLobeAngleFraction, specularLobeAngleFraction = cb.gLobeAngleFraction;
                const float2 pixelUv = float2(px + 0.5f, py + 0.5f) * cb.gRectSizeInv;
                if (cb.gHasHistoryConfidence) {  // :201-215
                    float relax = saturate(cb.gConfidenceDrivenRelaxationMultiplier * (1.0f - saturate(t.gIn_SpecConfidence->sampleLinear(pixelUv).x)));
                    float r = saturate(relax * cb.gConfidenceDrivenNormalEdgeStoppingRelaxation);
                    diffuseLobeAngleFractionForSimplifiedSpecularNormalWeight = lerp(diffuseLobeAngleFraction, 1.0f, r);
                    specularLobeAngleFraction = lerp(specularLobeAngleFraction, 1.0f, r);
                    r = saturate(relax * cb.gConfidenceDrivenLuminanceEdgeStoppingRelaxation);
                    specularLuminanceWeightRelaxation *= 1.0f - r;
                }
                float specularNormalWeightParamSimplified = GetNormalWeightParam2(1.0f, diffuseLobeAngleFractionForSimplifiedSpecularNormalWeight);
                float2 specularNormalWeightParams = GetNormalWeightParams_ATrous(centerRoughness, historyLength, specularReprojectionConfidence, cb.gNormalEdgeStoppingRelaxation,
                                                                                 specularLobeAngleFraction, cb.gSpecLobeAngleSlack);
                float3 centerV = -normalize(centerWorldPos);
                float centerDiffuseLuminance = Luminance(ctr.diff.xyz());
                float diffusePhiLIlluminationInv = 1.0f / max(1.0e-4f, cb.gDiffPhiLuminance * std::sqrt(centerDiffuseVar));
                float diffuseLuminanceWeightRelaxation = 1.0f;
                if (cb.gHasHistoryConfidence) {  // :239-251
                    float relax = saturate(cb.gConfidenceDrivenRelaxationMultiplier * (1.0f - saturate(t.gIn_DiffConfidence->sampleLinear(pixelUv).x)));
                    float r = saturate(relax * cb.gConfidenceDrivenNormalEdgeStoppingRelaxation);
                    diffuseLobeAngleFraction = lerp(diffuseLobeAngleFraction, 1.0f, r);
                    r = saturate(relax * cb.gConfidenceDrivenLuminanceEdgeStoppingRelaxation);
                    diffuseLuminanceWeightRelaxation = 1.0f - r;
                }
                float diffuseNormalWeightParam = GetNormalWeightParam2(1.0f, diffuseLobeAngleFraction);
                float depthThreshold = cb.gDepthThreshold * (cb.gOrthoMode == 0.0f ? centerViewZ : 1.0f);

                float sumWSpecular = 0.0f, sumWDiffuse = 0.0f;
                float4 sumSpecular = float4(0.0f), sumDiffuse = float4(0.0f);
                float3 sumSpecularSH = float3(0.0f), sumDiffuseSH = float3(0.0f);
                for (int j = -1; j <= 1; j++)
                    for (int i = -1; i <= 1; i++) {
                        int x = px + i, y = py + j;
                        bool isCenter = i == 0 && j == 0;
                        bool isInside = x >= 0 && y >= 0 && x < cb.gRectSize.x && y < cb.gRectSize.y;
                        float kernelW = isInside ? kGauss3[std::abs(i)] * kGauss3[std::abs(j)] : 0.0f;
                        Texel s = fetch(x, y);
                        float3 sampleNormal = s.nr.xyz();
                        float geometryW = GetPlaneDistanceWeight_Atrous(centerWorldPos, centerNormal, s.worldPos, depthThreshold);
                        geometryW *= kernelW;

                        float angles = Math::AcosApproxPositive(dot(centerNormal, sampleNormal));
                        float3 sampleV = -normalize(s.worldPos + cb.gRoughnessEdgeStoppingRelaxation * centerWorldPos);
                        float normalWSpecularSimplified = ComputeWeight(angles, specularNormalWeightParamSimplified, 0.0f);
                        float normalWSpecular = GetSpecularNormalWeight_ATrous(specularNormalWeightParams, centerNormal, sampleNormal, centerV, sampleV);
                        float roughnessWSpecular = ComputeWeight(s.nr.w, roughnessWeightParams.x, roughnessWeightParams.y);
                        float specularLuminanceW = std::fabs(centerSpecularLuminance - Luminance(s.spec.xyz())) * specularPhiLIlluminationInv;
                        specularLuminanceW = min(cb.gSpecMaxLuminanceRelativeDifference, specularLuminanceW);
                        specularLuminanceW *= specularLuminanceWeightRelaxation;
                        float wSpecular = geometryW * std::exp(-specularLuminanceW);
                        wSpecular *= cb.gRoughnessEdgeStoppingEnabled ? (normalWSpecular * roughnessWSpecular) : normalWSpecularSimplified;
                        wSpecular *= float(CompareMaterials(s.materialID, centerMaterialID, cb.gSpecMinMaterial));
                        wSpecular = isCenter ? kernelW : wSpecular;
                        sumWSpecular += wSpecular;
                        sumSpecular += wSpecular * s.spec;
                        sumSpecularSH += wSpecular * s.specSh;

                        float normalWDiffuse = ComputeWeight(angles, diffuseNormalWeightParam, 0.0f);
                        float diffuseLuminanceW = std::fabs(centerDiffuseLuminance - Luminance(s.diff.xyz())) * diffusePhiLIlluminationInv;
                        diffuseLuminanceW = min(cb.gDiffMaxLuminanceRelativeDifference, diffuseLuminanceW);
                        diffuseLuminanceW *= diffuseLuminanceWeightRelaxation;
                        float wDiffuse = geometryW * normalWDiffuse * std::exp(-diffuseLuminanceW);
                        wDiffuse *= float(CompareMaterials(s.materialID, centerMaterialID, cb.gDiffMinMaterial));
                        wDiffuse = isCenter ? kernelW : wDiffuse;
                        sumWDiffuse += wDiffuse;
                        sumDiffuse += wDiffuse * s.diff;
                        sumDiffuseSH += wDiffuse * s.diffSh;
                    }
                sumWSpecular = max(sumWSpecular, 1e-6f);
                sumSpecular = sumSpecular / sumWSpecular;
                float sp1 = Luminance(sumSpecular.xyz());
                t.gOut_Spec_Variance->store(px, py, float4(sumSpecular.xyz(), max(0.0f, sumSpecular.w - sp1 * sp1)));
                t.gOut_SpecSh->store(px, py, float4(sumSpecularSH / sumWSpecular, 0.0f));
                sumWDiffuse = max(sumWDiffuse, 1e-6f);
                sumDiffuse = sumDiffuse / sumWDiffuse;
                float dp1 = Luminance(sumDiffuse.xyz());
                t.gOut_Diff_Variance->store(px, py, float4(sumDiffuse.xyz(), max(0.0f, sumDiffuse.w - dp1 * dp1)));
                t.gOut_DiffSh->store(px, py, float4(sumDiffuseSH / sumWDiffuse, 0.0f));
            } else {
                // spatial variance estimation over 5x5
                float sumWS = 0.0f, sumS1 = 0.0f, sumS2 = 0.0f, sumWD = 0.0f, sumD1 = 0.0f, sumD2 = 0.0f;
                float3 sumS = float3(0.0f), sumD = float3(0.0f), sumSSH = float3(0.0f), sumDSH = float3(0.0f);
                float diffuseNormalWeightParam = GetNormalWeightParam2(1.0f, cb.gLobeAngleFraction);
                for (int cx = -2; cx <= 2; cx++)
                    for (int cy = -2; cy <= 2; cy++) {
                        Texel s = fetch(px + cx, py + cy);
                        float angle = Math::AcosApproxPositive(dot(centerNormal, s.nr.xyz()));
                        float normalW = ComputeWeight(angle, diffuseNormalWeightParam, 0.0f);
                        float specularW = normalW * float(CompareMaterials(s.materialID, centerMaterialID, cb.gSpecMinMaterial));
                        sumWS += specularW;
                        sumS += s.spec.xyz() * specularW;
                        sumS1 += Luminance(s.spec.xyz()) * specularW;
                        sumS2 += s.spec.w * specularW;
                        sumSSH += s.specSh * specularW;
                        float diffuseW = normalW * float(CompareMaterials(s.materialID, centerMaterialID, cb.gDiffMinMaterial));
                        sumWD += diffuseW;
                        sumD += s.diff.xyz() * diffuseW;
                        sumD1 += Luminance(s.diff.xyz()) * diffuseW;
                        sumD2 += s.diff.w * diffuseW;
                        sumDSH += s.diffSh * diffuseW;
                    }
                float boost = max(1.0f, 4.0f / (historyLength + 1.0f));
                sumWS = max(sumWS, 1e-6f);
                sumS = sumS / sumWS;
                sumS1 /= sumWS;
                sumS2 /= sumWS;
                t.gOut_Spec_Variance->store(px, py, float4(sumS, max(0.0f, sumS2 - sumS1 * sumS1) * boost));
                t.gOut_SpecSh->store(px, py, float4(sumSSH / sumWS, 0.0f));
                sumWD = max(sumWD, 1e-6f);
                sumD = sumD / sumWD;
                sumD1 /= sumWD;
                sumD2 /= sumWD;
                t.gOut_Diff_Variance->store(px, py, float4(sumD, max(0.0f, sumD2 - sumD1 * sumD1) * boost));
                t.gOut_DiffSh->store(px, py, float4(sumDSH / sumWD, 0.0f));
            }
        }
}

void atrous(const RelaxCB& cb, const AtTex& t, int gridW, int gridH) {
    Ctx c(cb);
    const float stepSize = (float)cb.gStepSize;
#pragma omp parallel for schedule(dynamic, 4)
    for (int py = 0; py < gridH * 16; py++)
        for (int px = 0; px < gridW * 16; px++) {
            if (t.gIn_Tiles->load(px >> 4, py >> 4).x != 0.0f || px >= cb.gRectSize.x || py >= cb.gRectSize.y) continue;
            float centerViewZ = c.UnpackViewZ(t.gIn_ViewZ->load(px, py).x);
            if (!c.IsInDenoisingRange(centerViewZ)) continue;

            float centerMaterialID;
            float4 centerNormalRoughness = unpackNR(*t.gIn_Normal_Roughness, px, py, centerMaterialID);
            float3 centerNormal = centerNormalRoughness.xyz();
            float centerRoughness = centerNormalRoughness.w;
            float historyLength = 255.0f * t.gIn_HistoryLength->load(px, py).x;

            float diffuseLobeAngleFraction = 1.0f / std::sqrt(stepSize);  // NRD_MODE == SH
            diffuseLobeAngleFraction = lerp(0.99f, diffuseLobeAngleFraction, saturate(historyLength / 5.0f));

            float4 centerSpecular = t.gIn_Spec_Variance->load(px, py);
            float centerSpecularLuminance = Luminance(centerSpecular.xyz());
            float specularPhiLIlluminationInv = 1.0f / max(1.0e-4f, cb.gSpecPhiLuminance * std::sqrt(centerSpecular.w));
            float2 roughnessWeightParams = GetRoughnessWeightParams(centerRoughness, cb.gRoughnessFraction);
            float specularReprojectionConfidence = t.gIn_SpecReprojectionConfidence->load(px, py).x;
            float specularLuminanceWeightRelaxation = 1.0f;
            if (cb.gStepSize <= 4) specularLuminanceWeightRelaxation = lerp(1.0f, specularReprojectionConfidence, cb.gLuminanceEdgeStoppingRelaxation);
            float diffuseLobeAngleFractionForSimplifiedSpecularNormalWeight = diffuseLobeAngleFraction, specularLobeAngleFraction = cb.gLobeAngleFraction;
            const float2 pixelUv = float2(px + 0.5f, py + 0.5f) * cb.gRectSizeInv;
            if (cb.gHasHistoryConfidence) {  // RELAX_Atrous.cs.hlsl:67-80
                float relax = saturate(cb.gConfidenceDrivenRelaxationMultiplier * (1.0f - saturate(t.gIn_SpecConfidence->sampleLinear(pixelUv).x)));
                float r = saturate(relax * cb.gConfidenceDrivenNormalEdgeStoppingRelaxation);
                diffuseLobeAngleFractionForSimplifiedSpecularNormalWeight = lerp(diffuseLobeAngleFraction, 1.0f, r);
                specularLobeAngleFraction = lerp(specularLobeAngleFraction, 1.0f, r);
                r = saturate(relax * cb.gConfidenceDrivenLuminanceEdgeStoppingRelaxation);
                specularLuminanceWeightRelaxation *= 1.0f - r;
            }
            float specularNormalWeightParamSimplified = GetNormalWeightParam2(1.0f, diffuseLobeAngleFractionForSimplifiedSpecularNormalWeight);
            float2 specularNormalWeightParams = GetNormalWeightParams_ATrous(centerRoughness, historyLength, specularReprojectionConfidence, cb.gNormalEdgeStoppingRelaxation,
                                                                             specularLobeAngleFraction, cb.gSpecLobeAngleSlack);
            float sumWSpecular = 0.44198f * 0.44198f;
            float4 sumSpecular = centerSpecular * float4(sumWSpecular, sumWSpecular, sumWSpecular, sumWSpecular * sumWSpecular);
            float3 sumSpecularSH = t.gIn_SpecSh->load(px, py).xyz() * sumWSpecular;

            float4 centerDiffuse = t.gIn_Diff_Variance->load(px, py);
            float centerDiffuseLuminance = Luminance(centerDiffuse.xyz());
            float diffusePhiLIlluminationInv = 1.0f / max(1.0e-4f, cb.gDiffPhiLuminance * std::sqrt(centerDiffuse.w));
            float diffuseLuminanceWeightRelaxation = 1.0f;
            if (cb.gHasHistoryConfidence) {  // :107-119
                float relax = saturate(cb.gConfidenceDrivenRelaxationMultiplier * (1.0f - saturate(t.gIn_DiffConfidence->sampleLinear(pixelUv).x)));
                float r = saturate(relax * cb.gConfidenceDrivenNormalEdgeStoppingRelaxation);
                diffuseLobeAngleFraction = lerp(diffuseLobeAngleFraction, 1.0f, r);
                r = saturate(relax * cb.gConfidenceDrivenLuminanceEdgeStoppingRelaxation);
                diffuseLuminanceWeightRelaxation = 1.0f - r;
            }
            float diffuseNormalWeightParam = GetNormalWeightParam2(1.0f, diffuseLobeAngleFraction);
            float sumWDiffuse = 0.44198f * 0.44198f;
            float4 sumDiffuse = centerDiffuse * float4(sumWDiffuse, sumWDiffuse, sumWDiffuse, sumWDiffuse * sumWDiffuse);
            float3 sumDiffuseSH = t.gIn_DiffSh->load(px, py).xyz() * sumWDiffuse;

            float3 centerWorldPos = c.GetCurrentWorldPosFromPixelPos(px, py, centerViewZ);
            float3 centerV = -normalize(centerWorldPos);
            float depthThreshold = cb.gDepthThreshold * (cb.gOrthoMode == 0.0f ? centerViewZ : 1.0f);

            int offX = 0, offY = 0;
            if (cb.gStepSize > 4) {
                RngHash rng;
                rng.Initialize((uint32_t)px, (uint32_t)py, cb.gFrameIndex);
                float2 r = rng.GetFloat2();
                offX = (int)(stepSize * 0.5f * (r.x - 0.5f));
                offY = (int)(stepSize * 0.5f * (r.y - 0.5f));
            }
            for (int j = -1; j <= 1; j++)
                for (int i = -1; i <= 1; i++) {
                    if (i == 0 && j == 0) continue;
                    int x = px + offX + i * (int)cb.gStepSize, y = py + offY + j * (int)cb.gStepSize;
                    bool isInside = x >= 0 && y >= 0 && x < cb.gRectSize.x && y < cb.gRectSize.y;
                    float kernelW = kGauss3[std::abs(i)] * kGauss3[std::abs(j)];

                    float sampleMaterialID;
                    float4 sampleNormalRoughness = unpackNR(*t.gIn_Normal_Roughness, x, y, sampleMaterialID);
                    float3 sampleNormal = sampleNormalRoughness.xyz();
                    float sampleViewZ = c.UnpackViewZ(t.gIn_ViewZ->load(x, y).x);
                    float3 sampleWorldPos = c.GetCurrentWorldPosFromPixelPos(x, y, sampleViewZ);
                    float geometryW = GetPlaneDistanceWeight_Atrous(centerWorldPos, centerNormal, sampleWorldPos, depthThreshold);
                    geometryW *= kernelW;
                    geometryW *= float(isInside && c.IsInDenoisingRange(sampleViewZ));

                    float3 sampleV = -normalize(sampleWorldPos + cb.gRoughnessEdgeStoppingRelaxation * centerWorldPos);
                    float angles = Math::AcosApproxPositive(dot(centerNormal, sampleNormal));
                    float normalWSpecularSimplified = ComputeWeight(angles, specularNormalWeightParamSimplified, 0.0f);
                    float normalWSpecular = GetSpecularNormalWeight_ATrous(specularNormalWeightParams, centerNormal, sampleNormal, centerV, sampleV);
                    float roughnessWSpecular = ComputeWeight(sampleNormalRoughness.w, roughnessWeightParams.x, roughnessWeightParams.y);
                    float wSpecular = geometryW * (cb.gRoughnessEdgeStoppingEnabled ? (normalWSpecular * roughnessWSpecular) : normalWSpecularSimplified);
                    wSpecular *= float(CompareMaterials(sampleMaterialID, centerMaterialID, cb.gSpecMinMaterial));
                    if (wSpecular > 1e-4f) {
                        float4 s = t.gIn_Spec_Variance->load(x, y);
                        float lw = std::fabs(centerSpecularLuminance - Luminance(s.xyz())) * specularPhiLIlluminationInv;
                        lw = min(cb.gSpecMaxLuminanceRelativeDifference, lw);
                        lw *= specularLuminanceWeightRelaxation;
                        wSpecular *= std::exp(-lw);
                        sumWSpecular += wSpecular;
                        sumSpecular += float4(wSpecular, wSpecular, wSpecular, wSpecular * wSpecular) * s;
                        sumSpecularSH += t.gIn_SpecSh->load(x, y).xyz() * wSpecular;
                    }

                    float normalWDiffuse = ComputeWeight(angles, diffuseNormalWeightParam, 0.0f);
                    float wDiffuse = geometryW * normalWDiffuse;
                    wDiffuse *= float(CompareMaterials(sampleMaterialID, centerMaterialID, cb.gDiffMinMaterial));
                    if (wDiffuse > 1e-4f) {
                        float4 s = t.gIn_Diff_Variance->load(x, y);
                        float lw = std::fabs(centerDiffuseLuminance - Luminance(s.xyz())) * diffusePhiLIlluminationInv;
                        lw = min(cb.gDiffMaxLuminanceRelativeDifference, lw);
                        lw *= diffuseLuminanceWeightRelaxation;
                        wDiffuse *= std::exp(-lw);
                        sumWDiffuse += wDiffuse;
                        sumDiffuse += float4(wDiffuse, wDiffuse, wDiffuse, wDiffuse * wDiffuse) * s;
                        sumDiffuseSH += t.gIn_DiffSh->load(x, y).xyz() * wDiffuse;
                    }
                }
            float currHistoryLength = max(historyLength - 1.0f, 0.0f);
            float4 filteredSpecular = sumSpecular / float4(sumWSpecular, sumWSpecular, sumWSpecular, sumWSpecular * sumWSpecular);
            if (cb.gIsLastPass == 1) filteredSpecular = float4(_NRD_LinearToYCoCg(filteredSpecular.xyz()), currHistoryLength);
            t.gOut_SpecSh->store(px, py, float4(sumSpecularSH / sumWSpecular, 0.0f));
            t.gOut_Spec_Variance->store(px, py, filteredSpecular);
            float4 filteredDiffuse = sumDiffuse / float4(sumWDiffuse, sumWDiffuse, sumWDiffuse, sumWDiffuse * sumWDiffuse);
            if (cb.gIsLastPass == 1) filteredDiffuse = float4(_NRD_LinearToYCoCg(filteredDiffuse.xyz()), currHistoryLength);
            t.gOut_DiffSh->store(px, py, float4(sumDiffuseSH / sumWDiffuse, 0.0f));
            t.gOut_Diff_Variance->store(px, py, filteredDiffuse);
        }
}

// RELAX_HitDistReconstruction.cs.hlsl:21-158 (NRD_SIGNAL = BOTH; one permutation serves SH and RADIANCE: only the .w of the SH0 / radiance
// textures is touched). `border` = 1 (3x3) or 2 (5x5); the shared-memory tile holds f( clamp( pos, 0, rectSize - 1 ) ). 8x8 groups.
void hitDistReconstruction(const RelaxCB& cb, const Tex& gIn_Tiles, const Tex& gIn_Normal_Roughness, const Tex& gIn_ViewZ, const Tex& gIn_Spec, const Tex& gIn_Diff, Tex& gOut_Spec,
                           Tex& gOut_Diff, int gridW, int gridH, int border) {
    Ctx c(cb);
    auto clampX = [&](int x) { return std::min(std::max(x, 0), cb.gRectSize.x - 1); };
    auto clampY = [&](int y) { return std::min(std::max(y, 0), cb.gRectSize.y - 1); };
#pragma omp parallel for schedule(dynamic, 4)
    for (int py = 0; py < gridH * 8; py++)
        for (int px = 0; px < gridW * 8; px++) {
            if (gIn_Tiles.load(px >> 4, py >> 4).x != 0.0f || px >= cb.gRectSize.x || py >= cb.gRectSize.y) continue;
            float centerViewZ = c.UnpackViewZ(gIn_ViewZ.load(px, py).x);
            if (!c.IsInDenoisingRange(centerViewZ)) continue;
            float2 pixelUv = float2(px + 0.5f, py + 0.5f) * cb.gRectSizeInv;
            float4 normalAndRoughness = unpackNR(gIn_Normal_Roughness, px, py);
            float3 centerNormal = normalAndRoughness.xyz();
            float centerRoughness = normalAndRoughness.w;

            float4 centerSpec = gIn_Spec.load(px, py), centerDiff = gIn_Diff.load(px, py);
            float centerSpecularHitDist = centerSpec.w, centerDiffuseHitDist = centerDiff.w;
            float2 relaxedRoughnessWeightParams = GetRelaxedRoughnessWeightParams(centerRoughness * centerRoughness);
            float specularNormalWeightParam = GetNormalWeightParam(1.0f, 1.0f, centerRoughness);
            float sumSpecularWeight = 1000.0f * float(centerSpecularHitDist != 0.0f);
            float sumSpecularHitDist = centerSpecularHitDist * sumSpecularWeight;
            float diffuseNormalWeightParam = GetNormalWeightParam(1.0f, 1.0f);
            float sumDiffuseWeight = 1000.0f * float(centerDiffuseHitDist != 0.0f);
            float sumDiffuseHitDist = centerDiffuseHitDist * sumDiffuseWeight;

            for (int dy = -border; dy <= border; dy++)
                for (int dx = -border; dx <= border; dx++) {
                    if (dx == 0 && dy == 0) continue;
                    int sx = clampX(px + dx), sy = clampY(py + dy);
                    float3 sampleNormal = unpackNR(gIn_Normal_Roughness, sx, sy).xyz();
                    float sampleViewZ = c.UnpackViewZ(gIn_ViewZ.load(sx, sy).x);
                    float angle = Math::AcosApproxPositive(dot(centerNormal, sampleNormal));
                    float2 o = float2((float)dx, (float)dy);

                    float w = IsInScreenNearest(pixelUv + o * cb.gRectSizeInv);
                    w *= float(c.IsInDenoisingRange(sampleViewZ));
                    w *= GetGaussianWeight(length(o) * 0.5f);
                    w *= GetBilateralWeight(sampleViewZ, centerViewZ);

                    float specularWeight = w;
                    specularWeight *= ComputeExponentialWeight(angle, specularNormalWeightParam, 0.0f);
                    // ( the shader weights by the CENTER roughness here, :121 )
                    specularWeight *= ComputeExponentialWeight(normalAndRoughness.w * normalAndRoughness.w, relaxedRoughnessWeightParams.x, relaxedRoughnessWeightParams.y);
                    float sampleSpecularHitDist = gIn_Spec.load(sx, sy).w;
                    if (specularWeight == 0.0f) sampleSpecularHitDist = 0.0f;
                    specularWeight *= float(sampleSpecularHitDist != 0.0f);
                    sumSpecularHitDist += sampleSpecularHitDist * specularWeight;
                    sumSpecularWeight += specularWeight;

                    float diffuseWeight = w;
                    diffuseWeight *= ComputeExponentialWeight(angle, diffuseNormalWeightParam, 0.0f);
                    float sampleDiffuseHitDist = gIn_Diff.load(sx, sy).w;
                    if (diffuseWeight == 0.0f) sampleDiffuseHitDist = 0.0f;
                    diffuseWeight *= float(sampleDiffuseHitDist != 0.0f);
                    sumDiffuseHitDist += diffuseWeight == 0.0f ? 0.0f : sampleDiffuseHitDist * diffuseWeight;
                    sumDiffuseWeight += diffuseWeight;
                }
            sumSpecularHitDist /= max(sumSpecularWeight, 1e-6f);
            gOut_Spec.store(px, py, float4(centerSpec.xyz(), sumSpecularHitDist));
            sumDiffuseHitDist /= max(sumDiffuseWeight, 1e-6f);
            gOut_Diff.store(px, py, float4(centerDiff.xyz(), sumDiffuseHitDist));
        }
}

// RELAX_SplitScreen.cs.hlsl:21-62 (NRD_SIGNAL = BOTH, NRD_MODE = SH)
void splitScreen(const RelaxCB& cb, const Tex& gIn_ViewZ, const Tex& gIn_Diff, const Tex& gIn_Spec, const Tex& gIn_DiffSh, const Tex& gIn_SpecSh, Tex& gOut_Diff, Tex& gOut_Spec,
                 Tex& gOut_DiffSh, Tex& gOut_SpecSh, int gridW, int gridH) {
    Ctx c(cb);
    for (int py = 0; py < gridH * 16; py++)
        for (int px = 0; px < gridW * 8; px++) {
            float2 pixelUv = float2(px + 0.5f, py + 0.5f) * cb.gRectSizeInv;
            if (pixelUv.x > cb.gSplitScreen || px >= cb.gRectSize.x || py >= cb.gRectSize.y) continue;
            float inRange = float(c.IsInDenoisingRange(c.UnpackViewZ(gIn_ViewZ.load(px, py).x)));
            int dx = px >> (cb.gDiffCheckerboard != 2 ? 1 : 0), sx = px >> (cb.gSpecCheckerboard != 2 ? 1 : 0);
            float4 diff = gIn_Diff.load(dx, py), spec = gIn_Spec.load(sx, py);
            diff = float4(_NRD_LinearToYCoCg(diff.xyz()), diff.w);
            spec = float4(_NRD_LinearToYCoCg(spec.xyz()), spec.w);
            gOut_Diff.store(px, py, diff * inRange);
            gOut_Spec.store(px, py, spec * inRange);
            gOut_DiffSh.store(px, py, float4(gIn_DiffSh.load(dx, py).xyz() * inRange, 0.0f));
            gOut_SpecSh.store(px, py, float4(gIn_SpecSh.load(sx, py).xyz() * inRange, 0.0f));
        }
}

}  // namespace

// returns 0 on success, 1 unknown shader, 2 bad arguments (same contract as nrd_oracle_dispatch)
int relaxDispatch(const std::string& id, const void* constants, uint32_t cbSize, Tex* t, uint32_t n, int gridW, int gridH) {
    if (cbSize != sizeof(RelaxCB)) return 2;
    const RelaxCB& cb = *(const RelaxCB*)constants;
    const std::string sig = "|NRD_SIGNAL=BOTH|NRD_MODE=SH";
    if (id == "RELAX_ClassifyTiles.cs.hlsl") {
        if (n != 2) return 2;
        classifyTiles(cb, t[0], t[1], gridW, gridH);
        return 0;
    }
    if (id == "RELAX_HitDistReconstruction.cs.hlsl|NRD_SIGNAL=BOTH|NRD_MODE=RADIANCE|MODE_5X5=0" || id == "RELAX_HitDistReconstruction.cs.hlsl|NRD_SIGNAL=BOTH|NRD_MODE=RADIANCE|MODE_5X5=1") {
        if (n != 7) return 2;
        hitDistReconstruction(cb, t[0], t[1], t[2], t[3], t[4], t[5], t[6], gridW, gridH, id.back() == '1' ? 2 : 1);
        return 0;
    }
    if (id == "RELAX_SplitScreen.cs.hlsl" + sig) {
        if (n != 9) return 2;
        splitScreen(cb, t[0], t[1], t[2], t[3], t[4], t[5], t[6], t[7], t[8], gridW, gridH);
        return 0;
    }
    if (id == "RELAX_PrePass.cs.hlsl" + sig) {
        if (n != 11) return 2;
        prePass(cb, t[0], t[1], t[2], t[3], t[4], t[5], t[6], t[7], t[8], t[9], t[10], gridW, gridH);
        return 0;
    }
    if (id == "RELAX_TemporalAccumulation.cs.hlsl" + sig) {
        if (n != 35) return 2;
        TaTex a = {&t[0], &t[1], &t[2], &t[3], &t[4], &t[5], &t[6], &t[7], &t[8], &t[9], &t[10], &t[11], &t[12], &t[13], &t[14], &t[15], &t[16], &t[17], &t[18], &t[19],
                   &t[20], &t[21], &t[22], &t[23], &t[24], &t[25], &t[26], &t[27], &t[28], &t[29], &t[30], &t[31], &t[32], &t[33], &t[34]};
        temporalAccumulation(cb, a, gridW, gridH);
        return 0;
    }
    if (id == "RELAX_HistoryFix.cs.hlsl" + sig) {
        if (n != 12) return 2;
        historyFix(cb, t[0], t[1], t[2], t[3], t[4], t[5], t[6], t[7], t[8], t[9], t[10], t[11], gridW, gridH);
        return 0;
    }
    if (id == "RELAX_HistoryClamping.cs.hlsl" + sig) {
        if (n != 22) return 2;
        HcTex a = {&t[0], &t[1], &t[2], &t[3], &t[4], &t[5], &t[6], &t[7], &t[8], &t[9], &t[10], &t[11], &t[12], &t[13], &t[14], &t[15], &t[16], &t[17], &t[18], &t[19], &t[20], &t[21]};
        historyClamping(cb, a, gridW, gridH);
        return 0;
    }
    if (id == "RELAX_Copy.cs.hlsl" + sig) {
        if (n != 4) return 2;
        copy(t[0], t[1], t[2], t[3], gridW, gridH);
        return 0;
    }
    if (id == "RELAX_AntiFirefly.cs.hlsl" + sig) {
        if (n != 7) return 2;
        antiFirefly(cb, t[0], t[1], t[2], t[3], t[4], t[5], t[6], gridW, gridH);
        return 0;
    }
    if (id == "RELAX_AtrousSmem.cs.hlsl" + sig) {
        if (n != 18) return 2;
        AtTex a = {&t[0], &t[1], &t[2], &t[3], &t[4], &t[5], &t[6], &t[7], &t[8], &t[9], &t[10], &t[11], &t[12], &t[13], &t[14], &t[15], &t[16], &t[17]};
        atrousSmem(cb, a, gridW, gridH);
        return 0;
    }
    if (id == "RELAX_Atrous.cs.hlsl" + sig) {
        if (n != 15) return 2;
        AtTex a = {&t[0], &t[1], &t[2], &t[3], &t[4], &t[5], &t[6], &t[7], &t[8], &t[9], &t[10], &t[11], &t[12], nullptr, nullptr, nullptr, &t[13], &t[14]};
        atrous(cb, a, gridW, gridH);
        return 0;
    }
    return 1;
}

}  // namespace orc
