// TEST INFRASTRUCTURE — CPU oracle (see hlsl_like.h). A hand restatement of the HLSL.
// PINNED: bit-identical, dispatch by dispatch, to the reference's own shaders compiled as C++
// (oracle/_ref/libnrd_refshaders.so, tests/test_oracle_vs_reference_shaders.py).
//
// REBLUR_DIFFUSE_SPECULAR passes (NRD_SIGNAL=BOTH, NRD_MODE=RADIANCE), one scalar function per compute shader:
//   ClassifyTiles          External/NRD/Shaders/REBLUR_ClassifyTiles.cs.hlsl:21-55
//   PrePass / Blur / PostBlur  REBLUR_PrePass.cs.hlsl:21-86, REBLUR_Blur.cs.hlsl:21-92, REBLUR_PostBlur.cs.hlsl:21-95,
//                          shared body REBLUR_Common_SpatialFilter.hlsli:59-336
//   TemporalAccumulation   REBLUR_TemporalAccumulation.cs.hlsl:68-995
//   HistoryFix             REBLUR_HistoryFix.cs.hlsl:44-506
//   TemporalStabilization  REBLUR_TemporalStabilization.cs.hlsl:41-318
//   Clear                  Clear.cs.hlsl:18-24
// Threads are replayed one pixel at a time over the dispatch grid; group-shared tiles become clamped texture reads
// and SM 6.0 quad ops (4 consecutive lanes of the flattened 8x16 group = x^1 / x^2 in the same row) are replayed
// by evaluating the neighbour lanes.
#include <omp.h>

#include "reblur_passes.h"

namespace orc {

namespace {

inline float4 unpackNR(const Tex& t, int x, int y, float& materialID) { return NRD_FrontEnd_UnpackNormalAndRoughness(t.load(x, y), materialID); }
inline float4 unpackNR(const Tex& t, int x, int y) { float m; return unpackNR(t, x, y, m); }
inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

}  // namespace

// ------------------------------------------------------------------------------------------------------------
void reblurClassifyTiles(const ReblurCB& cb, const Tex& gIn_ViewZ, Tex& gOut_Tiles, int gridW, int gridH) {
    ReblurCtx c(cb);
#pragma omp parallel for schedule(static)
    for (int ty = 0; ty < gridH; ty++)
        for (int tx = 0; tx < gridW; tx++) {
            int sum = 0;
            for (int j = 0; j < 16; j++)
                for (int i = 0; i < 16; i++) {
                    float viewZ = c.UnpackViewZ(gIn_ViewZ.load(tx * 16 + i, ty * 16 + j).x);
                    sum += !c.IsInDenoisingRange(viewZ) ? 1 : 0;
                }
            gOut_Tiles.store(tx, ty, float4(sum == 256 ? 1.0f : 0.0f));
        }
}

// ------------------------------------------------------------------------------------------------------------
// Spatial passes
// ------------------------------------------------------------------------------------------------------------
enum SpatialPass { PRE_PASS = 0, BLUR = 1, POST_BLUR = 2 };
enum Lobe { DIFF = 0, SPEC = 1 };

struct SpatialCommon {
    int px, py;
    float viewZ, materialID, roughness, NoV, frustumSize;
    float3 N, Nv, Xv, Vv;
    float2 pixelUv, nonLinearAccumSpeed, data1;
    float4 rotator;
    // checkerboard resolve of the pre-pass ( REBLUR_PrePass.cs.hlsl:52-67 ): parity of the pixel, half-res x of the left / right neighbours, their weights
    uint32_t checkerboard = 0;
    int checkerboardX0 = 0, checkerboardX1 = 0;
    float2 wc = float2(0.0f);
};

// REBLUR_Common_SpatialFilter.hlsli, one lobe. Checkerboard ( gDiffCheckerboard / gSpecCheckerboard != 2 ) only concerns the pre-pass:
// its input is half width, taps move to a pixel that was traced this frame, and pixels without data are resolved from their row neighbours.
template <int PASS, int LOBE>
static void spatialFilter(const ReblurCtx& c, const SpatialCommon& s, const Tex& gIn_ViewZ, const Tex& gIn_Normal_Roughness, const Tex& INPUT, Tex& OUTPUT,
                          Tex* gOut_SpecHitDistForTracking, Tex* OUTPUT_COPY, bool temporalStabilization, bool robustMirrorTest) {
    const ReblurCB& cb = c.cb;
    const float ROUGHNESS = LOBE == DIFF ? 1.0f : s.roughness;
    const float NON_LINEAR_ACCUM_SPEED = LOBE == DIFF ? s.nonLinearAccumSpeed.x : s.nonLinearAccumSpeed.y;
    const float MIN_MATERIAL = LOBE == DIFF ? cb.gDiffMinMaterial : cb.gSpecMinMaterial;
    const float MAX_BLUR_RADIUS = PASS == PRE_PASS ? (LOBE == DIFF ? cb.gDiffPrepassBlurRadius : cb.gSpecPrepassBlurRadius) : cb.gMaxBlurRadius;
    const bool USE_SCREEN_SPACE = LOBE == DIFF;  // REBLUR_USE_SCREEN_SPACE_SAMPLING_FOR_DIFFUSE = 1, ..._FOR_SPECULAR = 0

    const uint32_t CHECKERBOARD = PASS == PRE_PASS ? (LOBE == DIFF ? cb.gDiffCheckerboard : cb.gSpecCheckerboard) : 2u;

    float sum = 1.0f;
    float4 result = INPUT.load(CHECKERBOARD == 2 ? s.px : s.px >> 1, s.py);
    RngHash rng;
    if (CHECKERBOARD != 2 && s.checkerboard != CHECKERBOARD) {
        sum = 0.0f;
        result = float4(0.0f);
    }

    bool runFilter = PASS != PRE_PASS || MAX_BLUR_RADIUS != 0.0f;
    if (runFilter) {
        if (PASS == PRE_PASS && LOBE == SPEC) rng.Initialize((uint32_t)s.px, (uint32_t)s.py, cb.gFrameIndex);

        const float radiusScale = PASS == POST_BLUR ? 2.0f : 1.0f;
        const float fractionScale = PASS == PRE_PASS ? 2.0f : (PASS == BLUR ? 1.0f : 0.5f);

        // Hit distance factor
        float4 Dv = ImportanceSampling::GetSpecularDominantDirectionG2(s.Nv, s.Vv, ROUGHNESS);
        float NoD = std::fabs(dot(s.Nv, Dv.xyz()));
        float smc = GetSpecMagicCurve(ROUGHNESS, 0.5f);

        float hitDistScale = _REBLUR_GetHitDistanceNormalization(s.viewZ, cb.gHitDistSettings.xyz(), ROUGHNESS);
        float hitDist = result.w * hitDistScale;
        float hitDistFactor = GetHitDistFactor(hitDist, s.frustumSize);

        // Blur radius
        float areaFactor = PASS == PRE_PASS ? hitDistFactor : hitDistFactor * NON_LINEAR_ACCUM_SPEED;
        float blurRadius = radiusScale * Math::Sqrt01(areaFactor);
        blurRadius = saturate(blurRadius) * MAX_BLUR_RADIUS * smc;
        blurRadius = max(blurRadius, cb.gMinBlurRadius * smc);

        if (PASS == PRE_PASS && LOBE == SPEC) {
            float lobeTanHalfAngle = ImportanceSampling::GetSpecularLobeTanHalfAngle(ROUGHNESS, REBLUR_MAX_PERCENT_OF_LOBE_VOLUME_FOR_PRE_PASS);
            float worldLobeRadius = hitDist * NoD * lobeTanHalfAngle;
            float lobeRadius = worldLobeRadius / PixelRadiusToWorld(cb.gUnproject, cb.gOrthoMode, 1.0f, s.viewZ + hitDist * Dv.w);
            blurRadius = min(blurRadius, lobeRadius);
        }

        // Weights
        float2 geometryWeightParams = GetGeometryWeightParams(cb.gPlaneDistSensitivity, s.frustumSize, s.Xv, s.Nv);
        float normalWeightParam = GetNormalWeightParam(NON_LINEAR_ACCUM_SPEED, cb.gLobeAngleFraction, ROUGHNESS) / fractionScale;
        float2 roughnessWeightParams = GetRoughnessWeightParams(ROUGHNESS, cb.gRoughnessFraction * fractionScale);
        float2 hitDistanceWeightParams = GetHitDistanceWeightParams(result.w, NON_LINEAR_ACCUM_SPEED);
        float minHitDistWeight = cb.gMinHitDistanceWeight * fractionScale * smc;
        if (PASS != PRE_PASS) minHitDistWeight *= NON_LINEAR_ACCUM_SPEED;

        // Sampling space
        float4 scaledRotator;
        float3 Tv, Bv;
        if (PASS == PRE_PASS || USE_SCREEN_SPACE) {
            float2 skew = float2(1.0f);
            if (PASS != PRE_PASS && LOBE == DIFF) {
                skew = lerp(1.0f - abs(s.Nv.xy()), float2(1.0f), s.NoV);
                skew /= max(skew.x, skew.y);
            }
            skew *= cb.gRectSizeInv * blurRadius;
            scaledRotator = Geometry::ScaleRotator(s.rotator, skew);
        } else {
            float skewFactor;
            float3 bentDv;
            if (LOBE == DIFF) {
                skewFactor = 1.0f;
                bentDv = s.Nv;
            } else {
                float bentFactor = std::sqrt(hitDistFactor);
                skewFactor = lerp(0.25f + 0.75f * ROUGHNESS, 1.0f, NoD);
                skewFactor = lerp(skewFactor, 1.0f, NON_LINEAR_ACCUM_SPEED);
                skewFactor = lerp(1.0f, skewFactor, bentFactor);
                bentDv = normalize(lerp(s.Nv, Dv.xyz(), bentFactor));
            }
            float worldRadius = PixelRadiusToWorld(cb.gUnproject, cb.gOrthoMode, blurRadius, s.viewZ);
            ReblurCtx::GetKernelBasis(bentDv, s.Nv, Tv, Bv);
            Tv *= worldRadius * skewFactor;
            Bv *= worldRadius / skewFactor;
        }

        float hitDistForTracking = hitDist == 0.0f ? NRD_INF : hitDist;

        for (int n = 0; n < 8; n++) {
            float3 offset = g_Special8[n];

            float2 uv;
            if (PASS == PRE_PASS || USE_SCREEN_SPACE)
                uv = s.pixelUv + Geometry::RotateVector(scaledRotator, offset.xy());
            else
                uv = GetKernelSampleCoordinates(cb.gViewToClip, offset, s.Xv, Tv, Bv, s.rotator);

            float2 mirrorUv = MirrorUv(uv);
            // any( uv != mirrorUv ) is bit-fragile for in-screen taps ( 1 - ( 1 - uv ) re-rounds uv ); flag bit1 of
            // nrd_oracle_dispatch swaps in the intended "tap left the screen" test for the strict parity runs
            // ( "did mirroring MOVE the tap?": re-rounding moves an in-screen tap by <= 2^-24, leaving the screen by twice the overshoot )
            bool mirrored = robustMirrorTest ? (fabsf(uv.x - mirrorUv.x) > 1e-6f || fabsf(uv.y - mirrorUv.y) > 1e-6f) : (uv.x != mirrorUv.x || uv.y != mirrorUv.y);
            float w = mirrored ? 1.0f : GetGaussianWeight(offset.z);

            float2 posf = mirrorUv * cb.gRectSize;
            int2 pos = int2((int)posf.x, (int)posf.y);

            // Move to a "valid" pixel in checkerboard mode
            int checkerboardX = pos.x;
            if (CHECKERBOARD != 2) {
                int shift = ((n & 0x1) == 0) ? -1 : 1;
                pos.x += Sequence::CheckerBoard((uint32_t)pos.x, (uint32_t)pos.y, cb.gFrameIndex) != CHECKERBOARD ? shift : 0;
                checkerboardX = pos.x >> 1;
                w = (pos.x < 0 || pos.x > cb.gRectSizeMinusOne.x) ? 0.0f : w;
            }

            // Fetch data (PostBlur reads the copy of viewZ made by Blur — same texels)
            float zs = c.UnpackViewZ(gIn_ViewZ.load(pos).x);
            float3 Xvs = Geometry::ReconstructViewPosition(float2(pos.x + 0.5f, pos.y + 0.5f) * cb.gRectSizeInv, cb.gFrustum, zs, cb.gOrthoMode);

            float materialIDs;
            float4 Ns = unpackNR(gIn_Normal_Roughness, pos.x, pos.y, materialIDs);

            float angle = Math::AcosApproxPositive(dot(s.N, Ns.xyz()));
            float NoX = dot(s.Nv, Xvs);

            w *= CompareMaterials(s.materialID, materialIDs, MIN_MATERIAL) ? 1.0f : 0.0f;
            w *= ComputeWeight(angle, normalWeightParam, 0.0f);
            if (LOBE == SPEC) w *= ComputeWeight(Ns.w, roughnessWeightParams.x, roughnessWeightParams.y);
            w = c.ApplyGeometryWeightLast(w, zs, NoX, geometryWeightParams);

            float4 smp = INPUT.load(checkerboardX, pos.y);
            smp = w == 0.0f ? float4(0.0f) : smp;  // Denanify

            if (PASS == PRE_PASS && LOBE == SPEC) {
                float hs = smp.w * _REBLUR_GetHitDistanceNormalization(zs, cb.gHitDistSettings.xyz(), Ns.w);
                float geometryWeight = w * s.NoV * float(hs != 0.0f);
                if (rng.GetFloat() < geometryWeight) hitDistForTracking = min(hitDistForTracking, hs);

                w *= cb.gUsePrepassNotOnlyForSpecularMotionEstimation;

                float d = length(Xvs - s.Xv) + NRD_EPS;
                float t = hs / (d + hitDist);
                w *= lerp(saturate(t), 1.0f, Math::LinearStep(0.5f, 1.0f, ROUGHNESS));
            }

            w *= minHitDistWeight + ComputeExponentialWeight(smp.w, hitDistanceWeightParams.x, hitDistanceWeightParams.y);

            sum += w;
            result += smp * w;
        }

        float invSum = Math::PositiveRcp(sum);
        result *= invSum;

        if (PASS != PRE_PASS) result.w = hitDist / hitDistScale;

        if (PASS == PRE_PASS && LOBE == SPEC) gOut_SpecHitDistForTracking->store(s.px, s.py, float4(hitDistForTracking == NRD_INF ? 0.0f : hitDistForTracking));
    }

    // Checkerboard resolve ( if pre-pass failed )
    if (PASS == PRE_PASS && sum == 0.0f) {
        float4 s0 = INPUT.load(s.checkerboardX0, s.py);
        float4 s1 = INPUT.load(s.checkerboardX1, s.py);
        s0 = s.wc.x == 0.0f ? float4(0.0f) : s0;
        s1 = s.wc.y == 0.0f ? float4(0.0f) : s1;
        result = s0 * s.wc.x + s1 * s.wc.y;
    }

    OUTPUT.store(s.px, s.py, result);

    if (PASS == POST_BLUR && !temporalStabilization) {
        result.w = cb.gReturnHistoryLengthInsteadOfOcclusion ? (LOBE == DIFF ? s.data1.x : s.data1.y) : result.w;
        OUTPUT_COPY->store(s.px, s.py, result);
    }
}

// Center-pixel setup shared by the three spatial passes
static void spatialCenter(const ReblurCtx& c, SpatialCommon& s, const Tex& gIn_Normal_Roughness, float4 baseRotator) {
    const ReblurCB& cb = c.cb;
    float4 nr = unpackNR(gIn_Normal_Roughness, s.px, s.py, s.materialID);
    s.N = nr.xyz();
    s.Nv = Geometry::RotateVectorInverse(cb.gViewToWorld, s.N);
    s.roughness = nr.w;
    s.pixelUv = float2(s.px + 0.5f, s.py + 0.5f) * cb.gRectSizeInv;
    s.Xv = Geometry::ReconstructViewPosition(s.pixelUv, cb.gFrustum, s.viewZ, cb.gOrthoMode);
    s.Vv = c.GetViewVector(s.Xv, true);
    s.NoV = std::fabs(dot(s.Nv, s.Vv));
    s.frustumSize = GetFrustumSize(cb.gMinRectDimMulUnproject, cb.gOrthoMode, s.viewZ);
    s.rotator = baseRotator;  // *_ROTATOR_MODE = NRD_FRAME: the per-frame rotator is used as is (common:282-305)
}

// REBLUR_HitDistReconstruction.cs.hlsl:21-167 (NRD_SIGNAL = BOTH, RADIANCE, REBLUR_USE_DECOMPRESSED_HIT_DIST_IN_RECONSTRUCTION = 0,
// REBLUR_PERFORMANCE_MODE = 0). `border` = 1 (3x3) or 2 (5x5). The shared-memory tile holds f( clamp( pos, 0, rectSizeMinusOne ) ).
void reblurHitDistReconstruction(const ReblurCB& cb, const Tex& gIn_Tiles, const Tex& gIn_Normal_Roughness, const Tex& gIn_ViewZ, const Tex& gIn_Diff, const Tex& gIn_Spec,
                                 Tex& gOut_Diff, Tex& gOut_Spec, int gridW, int gridH, int border, int signal) {
    ReblurCtx c(cb, signal);
    auto clampPos = [&](int& x, int& y) {
        x = x < 0 ? 0 : (x > cb.gRectSizeMinusOne.x ? cb.gRectSizeMinusOne.x : x);
        y = y < 0 ? 0 : (y > cb.gRectSizeMinusOne.y ? cb.gRectSizeMinusOne.y : y);
    };
    auto hitDistViewZ = [&](int x, int y) {  // Preload( ): { diff hitDist, spec hitDist, viewZ }, hit distances zeroed outside the denoising range
        clampPos(x, y);
        float viewZ = c.UnpackViewZ(gIn_ViewZ.load(x, y).x);
        float2 hitDist = float2(0.0f);  // :33-48: the lobe the denoiser does not have stays 0
        if (c.hasDiff()) hitDist.x = gIn_Diff.load(x, y).w;
        if (c.hasSpec()) hitDist.y = gIn_Spec.load(x, y).w;
        if (!c.IsInDenoisingRange(viewZ)) hitDist = float2(0.0f);
        return float3(hitDist.x, hitDist.y, viewZ);
    };
    auto normalRoughness = [&](int x, int y) {
        clampPos(x, y);
        return NRD_FrontEnd_UnpackNormalAndRoughness(gIn_Normal_Roughness.load(x, y));
    };
#pragma omp parallel for schedule(dynamic, 4)
    for (int py = 0; py < gridH * 16; py++)
        for (int px = 0; px < gridW * 8; px++) {
            float isSky = gIn_Tiles.load(px >> 4, py >> 4).x;
            if (isSky != 0.0f || px > cb.gRectSizeMinusOne.x || py > cb.gRectSizeMinusOne.y) continue;
            float3 center = hitDistViewZ(px, py);
            if (!c.IsInDenoisingRange(center.z)) continue;

            float4 nr = normalRoughness(px, py);
            float3 N = nr.xyz();
            float roughness = nr.w;
            float2 pixelUv = float2(px + 0.5f, py + 0.5f) * cb.gRectSizeInv;
            float3 Xv = Geometry::ReconstructViewPosition(pixelUv, cb.gFrustum, center.z, cb.gOrthoMode);
            float3 Nv = Geometry::RotateVectorInverse(cb.gViewToWorld, N);
            float frustumSize = GetFrustumSize(cb.gMinRectDimMulUnproject, cb.gOrthoMode, center.z);
            float2 geometryWeightParams = GetGeometryWeightParams(cb.gPlaneDistSensitivity, frustumSize, Xv, Nv);
            float2 relaxedRoughnessWeightParams = GetRelaxedRoughnessWeightParams(roughness * roughness);
            float diffNormalWeightParam = GetNormalWeightParam(1.0f, 1.0f);
            float specNormalWeightParam = GetNormalWeightParam(1.0f, 1.0f, roughness);

            float2 sum = float2(center.x != 0.0f ? 1000.0f : 0.0f, center.y != 0.0f ? 1000.0f : 0.0f);
            float2 acc = float2(center.x, center.y) * sum;
            for (int j = 0; j <= border * 2; j++)
                for (int i = 0; i <= border * 2; i++) {
                    float2 o = float2(float(i - border), float(j - border));
                    if (o.x == 0.0f && o.y == 0.0f) continue;
                    int sx = px + i - border, sy = py + j - border;
                    float3 data = hitDistViewZ(sx, sy);

                    float2 uv = pixelUv + o * cb.gRectSizeInv;
                    float w = IsInScreenNearest(uv);
                    w *= GetGaussianWeight(length(o) * 0.5f);
                    float3 Xvs = Geometry::ReconstructViewPosition(uv, cb.gFrustum, data.z, cb.gOrthoMode);
                    float NoX = dot(Nv, Xvs);
                    w *= ComputeWeight(NoX, geometryWeightParams.x, geometryWeightParams.y);

                    float4 snr = normalRoughness(sx, sy);
                    float angle = Math::AcosApproxPositive(dot(N, snr.xyz()));
                    float2 ww = float2(w);
                    ww.x *= ComputeExponentialWeight(angle, diffNormalWeightParam, 0.0f);
                    ww.y *= ComputeExponentialWeight(angle, specNormalWeightParam, 0.0f);
                    ww.y *= ComputeExponentialWeight(snr.w * snr.w, relaxedRoughnessWeightParams.x, relaxedRoughnessWeightParams.y);
                    ww.x = data.x == 0.0f ? 0.0f : ww.x;
                    ww.y = data.y == 0.0f ? 0.0f : ww.y;

                    acc += float2(data.x, data.y) * ww;
                    sum += ww;
                }
            acc = acc / max(sum, float2(NRD_EPS));
            if (c.hasDiff()) gOut_Diff.store(px, py, float4(gIn_Diff.load(px, py).xyz(), acc.x));  // :150-166
            if (c.hasSpec()) gOut_Spec.store(px, py, float4(gIn_Spec.load(px, py).xyz(), acc.y));
        }
}

void reblurPrePass(const ReblurCB& cb, const Tex& gIn_Tiles, const Tex& gIn_Normal_Roughness, const Tex& gIn_ViewZ, const Tex& gIn_Diff, const Tex& gIn_Spec,
                   Tex& gOut_Diff, Tex& gOut_Spec, Tex& gOut_SpecHitDistForTracking, int gridW, int gridH, bool robust, int signal) {
    ReblurCtx c(cb, signal);
    const int W = gridW * 16, H = gridH * 16;
#pragma omp parallel for schedule(dynamic, 4)
    for (int py = 0; py < H; py++)
        for (int px = 0; px < W; px++) {
            float isSky = gIn_Tiles.load(px >> 4, py >> 4).x;
            if (isSky != 0.0f || px > cb.gRectSizeMinusOne.x || py > cb.gRectSizeMinusOne.y) continue;
            SpatialCommon s;
            s.px = px;
            s.py = py;
            s.viewZ = c.UnpackViewZ(gIn_ViewZ.load(px, py).x);
            if (!c.IsInDenoisingRange(s.viewZ)) continue;
            spatialCenter(c, s, gIn_Normal_Roughness, cb.gRotatorPre);
            s.nonLinearAccumSpeed = float2(1.0f / (1.0f + 10.0f));
            s.data1 = float2(0.0f);
            {  // Checkerboard resolve ( REBLUR_PrePass.cs.hlsl:52-67 )
                s.checkerboard = Sequence::CheckerBoard((uint32_t)px, (uint32_t)py, cb.gFrameIndex);
                int x0 = std::max(px - 1, 0), x1 = std::min(px + 1, cb.gRectSizeMinusOne.x);
                float viewZ0 = c.UnpackViewZ(gIn_ViewZ.load(x0, py).x), viewZ1 = c.UnpackViewZ(gIn_ViewZ.load(x1, py).x);
                float threshold = GetDisocclusionThreshold(NRD_DISOCCLUSION_THRESHOLD, s.frustumSize, s.NoV);
                float2 wc = float2(step(std::fabs(viewZ0 - s.viewZ), threshold), step(std::fabs(viewZ1 - s.viewZ), threshold));
                wc.x = (!c.IsInDenoisingRange(viewZ0) || px < 1) ? 0.0f : wc.x;
                wc.y = (!c.IsInDenoisingRange(viewZ1) || px >= cb.gRectSizeMinusOne.x) ? 0.0f : wc.y;
                wc *= Math::PositiveRcp(wc.x + wc.y);
                s.wc = wc;
                s.checkerboardX0 = x0 >> 1;
                s.checkerboardX1 = x1 >> 1;
            }
            if (c.hasDiff()) spatialFilter<PRE_PASS, DIFF>(c, s, gIn_ViewZ, gIn_Normal_Roughness, gIn_Diff, gOut_Diff, nullptr, nullptr, true, robust);  // :75-79
            if (c.hasSpec()) spatialFilter<PRE_PASS, SPEC>(c, s, gIn_ViewZ, gIn_Normal_Roughness, gIn_Spec, gOut_Spec, &gOut_SpecHitDistForTracking, nullptr, true, robust);  // :81-85
        }
}

// Value each lane holds when the quad exchange happens in Blur / PostBlur
static float2 blurNonLinearAccumSpeed(const ReblurCtx& c, const Tex& gIn_Data1, const Tex& viewZTex, int px, int py, float2* data1Out, float* viewZOut) {
    float4 d = gIn_Data1.load(px, py);
    float2 data1 = ReblurCtx::UnpackData1(float2(d.x, d.y), c.signal);
    float2 n = float2(c.GetAdvancedNonLinearAccumSpeed(data1.x), c.GetAdvancedNonLinearAccumSpeed(data1.y));
    float viewZ = c.UnpackViewZ(viewZTex.load(px, py).x);
    if (!c.IsInDenoisingRange(viewZ)) n = float2(0.0f);
    if (data1Out) *data1Out = data1;
    if (viewZOut) *viewZOut = viewZ;
    return n;
}

template <int PASS>
static void blurLike(const ReblurCB& cb, const Tex& gIn_Tiles, const Tex& gIn_Normal_Roughness, const Tex& gIn_ViewZ, const Tex& gIn_Data1, const Tex& gIn_Diff,
                     const Tex& gIn_Spec, Tex* gOut_ViewZ, Tex* gOut_Normal_Roughness, Tex* gOut_InternalData, Tex& gOut_Diff, Tex& gOut_Spec, Tex* gOut_DiffCopy,
                     Tex* gOut_SpecCopy, bool temporalStabilization, int gridW, int gridH, bool quads, bool robust, int signal) {
    ReblurCtx c(cb, signal);
    const int W = gridW * 8, H = gridH * 16;
#pragma omp parallel for schedule(dynamic, 4)
    for (int py = 0; py < H; py++)
        for (int px = 0; px < W; px++) {
            if (PASS == BLUR) {
                // Copy viewZ ( including sky! ) for the next pass and frame
                gOut_ViewZ->store(px, py, gIn_ViewZ.load(px, py));
            }
            float isSky = gIn_Tiles.load(px >> 4, py >> 4).x;
            if (isSky != 0.0f) continue;

            SpatialCommon s;
            s.px = px;
            s.py = py;
            s.nonLinearAccumSpeed = blurNonLinearAccumSpeed(c, gIn_Data1, gIn_ViewZ, px, py, &s.data1, &s.viewZ);
            if (quads) {
                float2 d10 = blurNonLinearAccumSpeed(c, gIn_Data1, gIn_ViewZ, px ^ 1, py, nullptr, nullptr);
                float2 d01 = blurNonLinearAccumSpeed(c, gIn_Data1, gIn_ViewZ, px ^ 2, py, nullptr, nullptr);
                float2 avg = (d10 + d01 + s.nonLinearAccumSpeed) / 3.0f;
                s.nonLinearAccumSpeed = min(s.nonLinearAccumSpeed, avg);
            }
            if (!c.IsInDenoisingRange(s.viewZ) || px > cb.gRectSizeMinusOne.x || py > cb.gRectSizeMinusOne.y) continue;

            spatialCenter(c, s, gIn_Normal_Roughness, PASS == BLUR ? cb.gRotator : cb.gRotatorPost);

            if (PASS == POST_BLUR) {
                gOut_Normal_Roughness->store(px, py, gIn_Normal_Roughness.load(px, py));  // same format: re-quantisation is the identity
                if (!temporalStabilization) gOut_InternalData->storeUint(px, py, c.PackInternalData(s.data1.x, s.data1.y, s.materialID));
            }
            if (c.hasDiff()) spatialFilter<PASS, DIFF>(c, s, gIn_ViewZ, gIn_Normal_Roughness, gIn_Diff, gOut_Diff, nullptr, gOut_DiffCopy, temporalStabilization, robust);
            if (c.hasSpec()) spatialFilter<PASS, SPEC>(c, s, gIn_ViewZ, gIn_Normal_Roughness, gIn_Spec, gOut_Spec, nullptr, gOut_SpecCopy, temporalStabilization, robust);
        }
}

void reblurBlur(const ReblurCB& cb, const Tex& gIn_Tiles, const Tex& gIn_Normal_Roughness, const Tex& gIn_ViewZ, const Tex& gIn_Data1, const Tex& gIn_Diff,
                const Tex& gIn_Spec, Tex& gOut_ViewZ, Tex& gOut_Diff, Tex& gOut_Spec, int gridW, int gridH, bool quads, bool robust, int signal) {
    blurLike<BLUR>(cb, gIn_Tiles, gIn_Normal_Roughness, gIn_ViewZ, gIn_Data1, gIn_Diff, gIn_Spec, &gOut_ViewZ, nullptr, nullptr, gOut_Diff, gOut_Spec, nullptr, nullptr,
                   true, gridW, gridH, quads, robust, signal);
}

void reblurPostBlur(const ReblurCB& cb, const Tex& gIn_Tiles, const Tex& gIn_Normal_Roughness, const Tex& gIn_Data1, const Tex& gIn_ViewZ, const Tex& gIn_Diff,
                    const Tex& gIn_Spec, Tex& gOut_Normal_Roughness, Tex& gOut_Diff, Tex& gOut_Spec, Tex* gOut_InternalData, Tex* gOut_DiffCopy, Tex* gOut_SpecCopy,
                    bool temporalStabilization, int gridW, int gridH, bool quads, bool robust, int signal) {
    blurLike<POST_BLUR>(cb, gIn_Tiles, gIn_Normal_Roughness, gIn_ViewZ, gIn_Data1, gIn_Diff, gIn_Spec, nullptr, &gOut_Normal_Roughness, gOut_InternalData, gOut_Diff,
                        gOut_Spec, gOut_DiffCopy, gOut_SpecCopy, temporalStabilization, gridW, gridH, quads, robust, signal);
}

// ------------------------------------------------------------------------------------------------------------
// Temporal accumulation
// ------------------------------------------------------------------------------------------------------------
namespace {

// s_Normal_HitDistForTracking[ y ][ x ] of the shader == Preload() evaluated at the clamped global position (TA:38-66)
float4 taPreload(const ReblurCtx& c, const TaTextures& t, int gx, int gy) {
    const ReblurCB& cb = c.cb;
    gx = clampi(gx, 0, cb.gRectSizeMinusOne.x);
    gy = clampi(gy, 0, cb.gRectSizeMinusOne.y);
    float3 N = unpackNR(*t.gIn_Normal_Roughness, gx, gy).xyz();
    if (!c.hasSpec()) return float4(N, 0.0f);  // TA:43-63: the tracking distance is a specular-only quantity
    float4 spec = t.gIn_Spec->load(gx, gy);
    float hitDist = cb.gSpecPrepassBlurRadius == 0.0f ? spec.w : t.gIn_SpecHitDistForTracking->load(gx, gy).x;
    float viewZ = c.UnpackViewZ(t.gIn_ViewZ->load(gx, gy).x);
    float hitDistForTracking = (hitDist == 0.0f || !c.IsInDenoisingRange(viewZ)) ? NRD_INF : hitDist;
    return float4(N, hitDistForTracking);
}

// TA:23-36 with REBLUR_USE_STF = 1 and R10G10B10A2 normals: returns the uv of ONE texel of the bilinear footprint
float2 stochasticBilinear(RngHash& rng, float2 uv, float2 texSize) {
    Filtering::Bilinear f = Filtering::GetBilinearFilter(uv, texSize);
    float2 rnd = rng.GetFloat2();
    f.origin += step(rnd, f.weights);
    return (f.origin + 0.5f) / texSize;
}

// gather of the 2x2 footprint whose top-left texel is (x0, y0), clamp addressing; order 00, 10, 01, 11 (== .wzxy)
float4 gather4(const Tex& t, int x0, int y0) {
    return float4(t.fetchClamped(x0, y0).x, t.fetchClamped(x0 + 1, y0).x, t.fetchClamped(x0, y0 + 1).x, t.fetchClamped(x0 + 1, y0 + 1).x);
}
float4 gather4Blue(const Tex& t, int x0, int y0) {
    return float4(t.fetchClamped(x0, y0).z, t.fetchClamped(x0 + 1, y0).z, t.fetchClamped(x0, y0 + 1).z, t.fetchClamped(x0 + 1, y0 + 1).z);
}
void gather4Uint(const Tex& t, int x0, int y0, uint32_t out[4]) {
    out[0] = t.fetchUintClamped(x0, y0);
    out[1] = t.fetchUintClamped(x0 + 1, y0);
    out[2] = t.fetchUintClamped(x0, y0 + 1);
    out[3] = t.fetchUintClamped(x0 + 1, y0 + 1);
}

}  // namespace

static void taPixel(const ReblurCtx& c, const TaTextures& t, int px, int py) {
    const ReblurCB& cb = c.cb;
    const float3 cameraDelta = cb.gCameraDelta.xyz();

    float viewZ = c.UnpackViewZ(t.gIn_ViewZ->load(px, py).x);
    if (!c.IsInDenoisingRange(viewZ)) return;

    // Current position
    float2 pixelUv = float2(px + 0.5f, py + 0.5f) * cb.gRectSizeInv;
    float3 Xv = Geometry::ReconstructViewPosition(pixelUv, cb.gFrustum, viewZ, cb.gOrthoMode);
    float3 X = Geometry::RotateVector(cb.gViewToWorld, Xv);

    // Hit distance for tracking (3x3 min), averaged normal (2x2, top-left quadrant of the 3x3 window)
    float3 Navg = float3(0.0f);
    float hitDistForTracking = NRD_INF;
    for (int j = 0; j <= 2; j++)
        for (int i = 0; i <= 2; i++) {
            float4 data = taPreload(c, t, px - 1 + i, py - 1 + j);
            if (i < 2 && j < 2) Navg += data.xyz() * 0.25f;
            hitDistForTracking = min(hitDistForTracking, data.w);
        }

    // Normal and roughness
    float materialID;
    float4 normalAndRoughness = unpackNR(*t.gIn_Normal_Roughness, px, py, materialID);
    float3 N = normalAndRoughness.xyz();
    float roughness = normalAndRoughness.w;

    float roughnessModified = Filtering::GetModifiedRoughnessFromNormalVariance(roughness, Navg);

    RngHash rng;
    rng.Initialize((uint32_t)px, (uint32_t)py, cb.gFrameIndex);

    const float hitDistNormalization = _REBLUR_GetHitDistanceNormalization(viewZ, cb.gHitDistSettings.xyz(), roughness);
    if (c.hasSpec()) {  // TA:123-142
        hitDistForTracking = hitDistForTracking == NRD_INF ? 0.0f : hitDistForTracking;
        hitDistForTracking *= cb.gSpecPrepassBlurRadius == 0.0f ? hitDistNormalization : 1.0f;
        t.gOut_SpecHitDistForTracking->store(px, py, float4(hitDistForTracking));
    }

    // Previous position and surface motion uv
    float4 mvRaw = t.gIn_Mv->load(px, py);
    float3 mv = mvRaw.xyz() * cb.gMvScale.xyz();
    float3 Xprev = X;
    float2 smbPixelUv = pixelUv + mv.xy();
    if (cb.gMvScale.w == 0.0f) {
        if (cb.gMvScale.z == 0.0f) mv.z = Geometry::AffineTransform(cb.gWorldToViewPrev, X).z - viewZ;
        float viewZprev = viewZ + mv.z;
        float3 Xvprevlocal = Geometry::ReconstructViewPosition(smbPixelUv, cb.gFrustumPrev, viewZprev, cb.gOrthoMode);
        Xprev = Geometry::RotateVectorInverse(cb.gWorldToViewPrev, Xvprevlocal) + cameraDelta;
    } else {
        Xprev += mv;
        smbPixelUv = Geometry::GetScreenUv(cb.gWorldToClipPrev, Xprev);
    }

    // Previous viewZ in the 4x4 CatRom footprint: four 2x2 gathers (TA:165-194)
    Filtering::CatmullRom smbCatromFilter = Filtering::GetCatmullRomFilter(smbPixelUv, cb.gRectSizePrev);
    // gather uv = origin * invSize, so the texel the sampler floors to is origin - 0.5 - 0.5 (+ offset)
    const int gx = (int)std::floor(smbCatromFilter.origin.x - 0.5f), gy = (int)std::floor(smbCatromFilter.origin.y - 0.5f);
    float4 smbViewZ0 = gather4(*t.gPrev_ViewZ, gx + 1, gy + 1);
    float4 smbViewZ1 = gather4(*t.gPrev_ViewZ, gx + 3, gy + 1);
    float4 smbViewZ2 = gather4(*t.gPrev_ViewZ, gx + 1, gy + 3);
    float4 smbViewZ3 = gather4(*t.gPrev_ViewZ, gx + 3, gy + 3);

    float3 prevViewZ0 = float3(c.UnpackViewZ(smbViewZ0.y), c.UnpackViewZ(smbViewZ0.z), c.UnpackViewZ(smbViewZ0.w));
    float3 prevViewZ1 = float3(c.UnpackViewZ(smbViewZ1.x), c.UnpackViewZ(smbViewZ1.z), c.UnpackViewZ(smbViewZ1.w));
    float3 prevViewZ2 = float3(c.UnpackViewZ(smbViewZ2.x), c.UnpackViewZ(smbViewZ2.y), c.UnpackViewZ(smbViewZ2.w));
    float3 prevViewZ3 = float3(c.UnpackViewZ(smbViewZ3.x), c.UnpackViewZ(smbViewZ3.y), c.UnpackViewZ(smbViewZ3.z));

    // Previous normal averaged over the 2x2 bilinear footprint
    Filtering::Bilinear smbBilinearFilter = Filtering::GetBilinearFilter(smbPixelUv, cb.gRectSizePrev);
    float smbNoN;
    float4 smbNoN2x2;
    {
        float3 Nt = Navg;  // yes, "Navg"
        int bx = (int)smbBilinearFilter.origin.x, by = (int)smbBilinearFilter.origin.y;
        float3 n00 = unpackNR(*t.gPrev_Normal_Roughness, bx, by).xyz();
        float3 n10 = unpackNR(*t.gPrev_Normal_Roughness, bx + 1, by).xyz();
        float3 n01 = unpackNR(*t.gPrev_Normal_Roughness, bx, by + 1).xyz();
        float3 n11 = unpackNR(*t.gPrev_Normal_Roughness, bx + 1, by + 1).xyz();
        smbNoN2x2 = float4(dot(n00, Nt), dot(n10, Nt), dot(n01, Nt), dot(n11, Nt));
        smbNoN = Filtering::ApplyBilinearFilter(smbNoN2x2.x, smbNoN2x2.y, smbNoN2x2.z, smbNoN2x2.w, smbBilinearFilter);
    }

    // Parallax
    float smbParallaxInPixels1 = ComputeParallaxInPixels(Xprev + cameraDelta, cb.gOrthoMode == 0.0f ? smbPixelUv : pixelUv, cb.gWorldToClipPrev, cb.gRectSize);
    float smbParallaxInPixels2 = ComputeParallaxInPixels(Xprev - cameraDelta, cb.gOrthoMode == 0.0f ? pixelUv : smbPixelUv, cb.gWorldToClip, cb.gRectSize);
    float smbParallaxInPixelsMax = max(smbParallaxInPixels1, smbParallaxInPixels2);
    float smbParallaxInPixelsMin = min(smbParallaxInPixels1, smbParallaxInPixels2);

    // Disocclusion: threshold
    float pixelSize = PixelRadiusToWorld(cb.gUnproject, cb.gOrthoMode, 1.0f, viewZ);
    float frustumSize = GetFrustumSize(cb.gMinRectDimMulUnproject, cb.gOrthoMode, viewZ);

    float disocclusionThresholdMix = 0.0f;
    if (materialID == cb.gStrandMaterialID) disocclusionThresholdMix = NRD_GetNormalizedStrandThickness(cb.gStrandThickness, pixelSize);
    if (cb.gHasDisocclusionThresholdMix) disocclusionThresholdMix = t.gIn_DisocclusionThresholdMix->load(px, py).x;

    float disocclusionThreshold = lerp(cb.gDisocclusionThreshold, cb.gDisocclusionThresholdAlternate, disocclusionThresholdMix);
    if (materialID == cb.gStrandMaterialID) {
        float mediumParallax = Math::SmoothStep01(smbParallaxInPixelsMax);
        disocclusionThreshold = lerp(0.25f, disocclusionThreshold, mediumParallax);
    }

    float smallParallax = Math::LinearStep(0.25f, 0.0f, smbParallaxInPixelsMax);
    float cosMaxAngle = std::cos(Math::DegToRad(89.0f)) - 0.25f * smallParallax;

    float3 V = c.GetViewVector(X);
    float NoV = std::fabs(dot(N, V));
    float NoVstrict = lerp(NoV, 1.0f, saturate(smbParallaxInPixelsMax / 30.0f));

    // Disocclusion
    float4 smbDisocclusionThreshold = float4(float(smbNoN2x2.x > cosMaxAngle), float(smbNoN2x2.y > cosMaxAngle), float(smbNoN2x2.z > cosMaxAngle), float(smbNoN2x2.w > cosMaxAngle));
    smbDisocclusionThreshold *= IsInScreenBilinear(smbBilinearFilter.origin, cb.gRectSizePrev);
    smbDisocclusionThreshold *= GetDisocclusionThreshold(disocclusionThreshold, frustumSize, NoVstrict);
    smbDisocclusionThreshold -= NRD_EPS;

    float3 Xvprev = Geometry::AffineTransform(cb.gWorldToViewPrev, Xprev);
    float3 smbPlaneDist0 = abs(prevViewZ0 - Xvprev.z);
    float3 smbPlaneDist1 = abs(prevViewZ1 - Xvprev.z);
    float3 smbPlaneDist2 = abs(prevViewZ2 - Xvprev.z);
    float3 smbPlaneDist3 = abs(prevViewZ3 - Xvprev.z);
    auto inRange3 = [&](float3 z) { return float3(float(c.IsInDenoisingRange(z.x)), float(c.IsInDenoisingRange(z.y)), float(c.IsInDenoisingRange(z.z))); };
    float3 smbOcclusion0 = step(smbPlaneDist0, float3(smbDisocclusionThreshold.x)) * inRange3(prevViewZ0);
    float3 smbOcclusion1 = step(smbPlaneDist1, float3(smbDisocclusionThreshold.y)) * inRange3(prevViewZ1);
    float3 smbOcclusion2 = step(smbPlaneDist2, float3(smbDisocclusionThreshold.z)) * inRange3(prevViewZ2);
    float3 smbOcclusion3 = step(smbPlaneDist3, float3(smbDisocclusionThreshold.w)) * inRange3(prevViewZ3);

    // Disocclusion: materialID
    uint32_t id0[4], id1[4], id2[4], id3[4];
    gather4Uint(*t.gPrev_InternalData, gx + 1, gy + 1, id0);
    gather4Uint(*t.gPrev_InternalData, gx + 3, gy + 1, id1);
    gather4Uint(*t.gPrev_InternalData, gx + 1, gy + 3, id2);
    gather4Uint(*t.gPrev_InternalData, gx + 3, gy + 3, id3);
    auto mat = [](uint32_t p) { return ReblurCtx::UnpackInternalData(p).z; };
    float3 smbMaterialID0 = float3(mat(id0[1]), mat(id0[2]), mat(id0[3]));
    float3 smbMaterialID1 = float3(mat(id1[0]), mat(id1[2]), mat(id1[3]));
    float3 smbMaterialID2 = float3(mat(id2[0]), mat(id2[1]), mat(id2[3]));
    float3 smbMaterialID3 = float3(mat(id3[0]), mat(id3[1]), mat(id3[2]));
    float minMaterialID = min(cb.gSpecMinMaterial, cb.gDiffMinMaterial);
    auto cmp3 = [&](float3 m) { return float3(float(CompareMaterials(materialID, m.x, minMaterialID)), float(CompareMaterials(materialID, m.y, minMaterialID)), float(CompareMaterials(materialID, m.z, minMaterialID))); };
    smbOcclusion0 *= cmp3(smbMaterialID0);
    smbOcclusion1 *= cmp3(smbMaterialID1);
    smbOcclusion2 *= cmp3(smbMaterialID2);
    smbOcclusion3 *= cmp3(smbMaterialID3);
    uint32_t smbInternalData[4] = {id0[3], id1[2], id2[1], id3[0]};

    // 2x2 occlusion weights
    float4 smbOcclusionWeights = Filtering::GetBilinearCustomWeights(smbBilinearFilter, float4(smbOcclusion0.z, smbOcclusion1.y, smbOcclusion2.y, smbOcclusion3.x));
    bool smbAllowCatRom = dot(smbOcclusion0 + smbOcclusion1 + smbOcclusion2 + smbOcclusion3, float3(1.0f)) > 11.5f;

    float fbits = smbOcclusion0.z * 1.0f;
    fbits += smbOcclusion1.y * 2.0f;
    fbits += smbOcclusion2.y * 4.0f;
    fbits += smbOcclusion3.x * 8.0f;

    // Accumulation speed
    float3 internalData00 = ReblurCtx::UnpackInternalData(smbInternalData[0]);
    float3 internalData10 = ReblurCtx::UnpackInternalData(smbInternalData[1]);
    float3 internalData01 = ReblurCtx::UnpackInternalData(smbInternalData[2]);
    float3 internalData11 = ReblurCtx::UnpackInternalData(smbInternalData[3]);
    float diffAccumSpeed = Filtering::ApplyBilinearCustomWeights(internalData00.x, internalData10.x, internalData01.x, internalData11.x, smbOcclusionWeights);
    float smbSpecAccumSpeed = Filtering::ApplyBilinearCustomWeights(internalData00.y, internalData10.y, internalData01.y, internalData11.y, smbOcclusionWeights);

    // Footprint quality
    float3 smbVprev = c.GetViewVectorPrev(Xprev, cameraDelta);
    float NoVprev = std::fabs(dot(N, smbVprev));
    float sizeQuality = (NoVprev + 1e-3f) / (NoV + 1e-3f);
    sizeQuality *= sizeQuality;
    sizeQuality = lerp(0.1f, 1.0f, saturate(sizeQuality));

    float smbFootprintQuality = Filtering::ApplyBilinearFilter(smbOcclusion0.z, smbOcclusion1.y, smbOcclusion2.y, smbOcclusion3.x, smbBilinearFilter);
    smbFootprintQuality = Math::Sqrt01(smbFootprintQuality);
    smbFootprintQuality *= sizeQuality;

    // ---------------------------------------------------------------------------------------------- Specular
    float specAccumSpeedCorrected = 0.0f, curvature = 0.0f, virtualHistoryAmount = 0.0f;  // TA:869-873 for a diffuse-only denoiser
    if (c.hasSpec()) {
        float smbSpecHistoryConfidence = smbFootprintQuality;
        if (cb.gHasHistoryConfidence) {
            float confidence = saturate(t.gIn_SpecConfidence->sampleLinear(smbPixelUv).x);
            smbSpecHistoryConfidence = min(smbSpecHistoryConfidence, confidence);
        }
        smbSpecAccumSpeed *= lerp(smbSpecHistoryConfidence, 1.0f, 1.0f / (1.0f + smbSpecAccumSpeed));

        const uint32_t checkerboard = Sequence::CheckerBoard((uint32_t)px, (uint32_t)py, cb.gFrameIndex);
        const bool specHasData = cb.gSpecCheckerboard == 2 || checkerboard == cb.gSpecCheckerboard;  // RADIANCE: the pre-pass has resolved the input already
        float4 spec = t.gIn_Spec->load(px, py);

        // Curvature estimation along predicted motion (TA:387-467)
        curvature = 0.0f;
        {
            float2 uvForZeroParallax = cb.gOrthoMode == 0.0f ? smbPixelUv : pixelUv;
            float2 deltaUv = uvForZeroParallax - Geometry::GetScreenUv(cb.gWorldToClipPrev, Xprev + cameraDelta);
            deltaUv *= cb.gRectSize;
            deltaUv /= max(smbParallaxInPixels1, 1.0f / 256.0f);

            float3 n10, x10;
            {
                float3 xv = Geometry::ReconstructViewPosition(pixelUv + float2(1, 0) * cb.gRectSizeInv, cb.gFrustum, 1.0f, cb.gOrthoMode);
                float3 x = Geometry::RotateVector(cb.gViewToWorld, xv);
                float3 v = c.GetViewVector(x);
                float3 o = cb.gOrthoMode == 0.0f ? float3(0.0f) : x;
                x10 = o + v * dot(X - o, N) / dot(N, v);
                n10 = taPreload(c, t, px + 1, py).xyz();
            }
            float3 n01, x01;
            {
                float3 xv = Geometry::ReconstructViewPosition(pixelUv + float2(0, 1) * cb.gRectSizeInv, cb.gFrustum, 1.0f, cb.gOrthoMode);
                float3 x = Geometry::RotateVector(cb.gViewToWorld, xv);
                float3 v = c.GetViewVector(x);
                float3 o = cb.gOrthoMode == 0.0f ? float3(0.0f) : x;
                x01 = o + v * dot(X - o, N) / dot(N, v);
                n01 = taPreload(c, t, px, py + 1).xyz();
            }

            float2 ww = abs(deltaUv) + 1.0f / 256.0f;
            ww /= ww.x + ww.y;

            float3 x = x10 * ww.x + x01 * ww.y;
            float3 n = normalize(n10 * ww.x + n01 * ww.y);

            float2 motionUvHigh = pixelUv + smbParallaxInPixelsMin * deltaUv * cb.gRectSizeInv;
            if (smbParallaxInPixelsMin > std::sqrt(2.0f) && IsInScreenNearest(motionUvHigh) != 0.0f) {
                // ClampUvToViewport, NRD_SUPPORTS_VIEWPORT_OFFSET = 0 (common:240)
                float2 uvScaled = min(motionUvHigh * cb.gResolutionScale, cb.gResolutionScale - 0.5f * cb.gResourceSizeInv);

                float zHigh = c.UnpackViewZ(t.gIn_ViewZ->sampleLinear(uvScaled).x);
                float3 xHigh = Geometry::ReconstructViewPosition(motionUvHigh, cb.gFrustum, zHigh, cb.gOrthoMode);
                xHigh = Geometry::RotateVector(cb.gViewToWorld, xHigh);

                float3 nHigh = NRD_FrontEnd_UnpackNormalAndRoughness(t.gIn_Normal_Roughness->sampleNearest(stochasticBilinear(rng, uvScaled, cb.gRectSize))).xyz();

                float2 geometryWeightParams = GetGeometryWeightParams(0.04f, frustumSize, X, N);
                float NoX = dot(N, xHigh);
                float w = c.ApplyGeometryWeightLast(1.0f, zHigh, NoX, geometryWeightParams);
                bool cmp = w > 0.5f;
                n = cmp ? nHigh : n;
                x = cmp ? xHigh : x;
            }

            float3 edge = x - X;
            float edgeLenSq = Math::LengthSquared(edge);
            curvature = dot(n - N, edge) * Math::PositiveRcp(edgeLenSq);

            if (curvature < 0.0f) {
                float2 uv1 = Geometry::GetScreenUv(cb.gWorldToClipPrev, GetXvirtual(hitDistForTracking, curvature, X, X, N, V, roughness));
                float2 uv2 = Geometry::GetScreenUv(cb.gWorldToClipPrev, X);
                float a = length((uv1 - uv2) * cb.gRectSize);
                curvature *= float(a < 5.0f * smbParallaxInPixelsMax + cb.gRectSizeInv.x);
            }
        }

        // Virtual motion - coordinates
        float3 Xvirtual = GetXvirtual(hitDistForTracking, curvature, X, Xprev, N, V, roughness);
        float XvirtualLength = length(Xvirtual);
        float hitDistanceToLobeSpreadInPixels = 1.0f / PixelRadiusToWorld(cb.gUnproject, cb.gOrthoMode, 1.0f, XvirtualLength);

        float2 vmbPixelUv = Geometry::GetScreenUv(cb.gWorldToClipPrev, Xvirtual);
        vmbPixelUv = materialID == cb.gCameraAttachedReflectionMaterialID ? smbPixelUv : vmbPixelUv;

        float2 vmbDelta = vmbPixelUv - smbPixelUv;
        float vmbPixelsTraveled = length(vmbDelta * cb.gRectSize);

        Filtering::Bilinear vmbBilinearFilter = Filtering::GetBilinearFilter(vmbPixelUv, cb.gRectSizePrev);
        // gather uv = (origin + 1) * invSize -> footprint top-left texel = origin
        const int vx = (int)vmbBilinearFilter.origin.x, vy = (int)vmbBilinearFilter.origin.y;

        // Virtual motion - confidence: roughness
        float virtualHistoryConfidence;
        float4 roughnessWeights;
        {
            float2 p = GetRelaxedRoughnessWeightParams(roughness * roughness, cb.gRoughnessFraction, REBLUR_ROUGHNESS_SENSITIVITY_IN_TA);
            float4 vmbRoughness = NRD_FrontEnd_UnpackRoughness(gather4Blue(*t.gPrev_Normal_Roughness, vx, vy));
            for (int i = 0; i < 4; i++) roughnessWeights[i] = ComputeNonExponentialWeight(vmbRoughness[i] * vmbRoughness[i], p.x, p.y);
            roughnessWeights = lerp(float4(1.0f), roughnessWeights, Math::SmoothStep01(vmbPixelsTraveled));
            virtualHistoryConfidence = Filtering::ApplyBilinearFilter(roughnessWeights.x, roughnessWeights.y, roughnessWeights.z, roughnessWeights.w, vmbBilinearFilter);
        }

        float4 vmbN;
        float4 vmbNoN2x2;
        float vmbNoN;
        {
            float3 Nt = N;  // yes, "N"
            float4 n00 = unpackNR(*t.gPrev_Normal_Roughness, vx, vy);
            float4 n10 = unpackNR(*t.gPrev_Normal_Roughness, vx + 1, vy);
            float4 n01 = unpackNR(*t.gPrev_Normal_Roughness, vx, vy + 1);
            float4 n11 = unpackNR(*t.gPrev_Normal_Roughness, vx + 1, vy + 1);
            vmbNoN2x2 = float4(dot(n00.xyz(), Nt), dot(n10.xyz(), Nt), dot(n01.xyz(), Nt), dot(n11.xyz(), Nt));
            vmbNoN = Filtering::ApplyBilinearFilter(vmbNoN2x2.x, vmbNoN2x2.y, vmbNoN2x2.z, vmbNoN2x2.w, vmbBilinearFilter);
            vmbN = Filtering::ApplyBilinearFilter(n00, n10, n01, n11, vmbBilinearFilter);
            float3 nn = _NRD_SafeNormalize(vmbN.xyz());
            vmbN = float4(nn, vmbN.w);
        }

        // Virtual motion - disocclusion
        float4 vmbOcclusionWeights;
        float vmbSpecAccumSpeed;
        bool vmbAllowCatRom;
        {
            float4 vmbOcclusionThreshold = float4(float(vmbNoN2x2.x > cosMaxAngle), float(vmbNoN2x2.y > cosMaxAngle), float(vmbNoN2x2.z > cosMaxAngle), float(vmbNoN2x2.w > cosMaxAngle));
            vmbOcclusionThreshold *= step(float4(0.5f), roughnessWeights);
            vmbOcclusionThreshold *= IsInScreenBilinear(vmbBilinearFilter.origin, cb.gRectSizePrev);
            vmbOcclusionThreshold *= disocclusionThreshold * frustumSize;
            vmbOcclusionThreshold *= lerp(0.1f, 1.0f, NoV);
            vmbOcclusionThreshold -= NRD_EPS;

            float4 vmbViewZraw = gather4(*t.gPrev_ViewZ, vx, vy);
            float4 vmbViewZ = float4(c.UnpackViewZ(vmbViewZraw.x), c.UnpackViewZ(vmbViewZraw.y), c.UnpackViewZ(vmbViewZraw.z), c.UnpackViewZ(vmbViewZraw.w));
            float3 vmbVv = Geometry::ReconstructViewPosition(vmbPixelUv, cb.gFrustumPrev, 1.0f);
            float3 Nv = Geometry::RotateVector(cb.gWorldToViewPrev, N);
            float NoXcurr = dot(N, Xprev - cameraDelta);
            float4 NoXprev = (Nv.x * vmbVv.x + Nv.y * vmbVv.y) * (cb.gOrthoMode == 0.0f ? vmbViewZ : float4(cb.gOrthoMode)) + Nv.z * vmbVv.z * vmbViewZ;
            float4 vmbPlaneDist = abs(NoXprev - NoXcurr);

            float4 inRange = float4(float(c.IsInDenoisingRange(vmbViewZ.x)), float(c.IsInDenoisingRange(vmbViewZ.y)), float(c.IsInDenoisingRange(vmbViewZ.z)), float(c.IsInDenoisingRange(vmbViewZ.w)));
            float4 vmbOcclusion = step(vmbPlaneDist, vmbOcclusionThreshold) * inRange;

            uint32_t vmbInternalData[4];
            gather4Uint(*t.gPrev_InternalData, vx, vy, vmbInternalData);
            float3 d00 = ReblurCtx::UnpackInternalData(vmbInternalData[0]);
            float3 d10 = ReblurCtx::UnpackInternalData(vmbInternalData[1]);
            float3 d01 = ReblurCtx::UnpackInternalData(vmbInternalData[2]);
            float3 d11 = ReblurCtx::UnpackInternalData(vmbInternalData[3]);

            vmbOcclusion.x *= float(CompareMaterials(materialID, d00.z, cb.gSpecMinMaterial));
            vmbOcclusion.y *= float(CompareMaterials(materialID, d10.z, cb.gSpecMinMaterial));
            vmbOcclusion.z *= float(CompareMaterials(materialID, d01.z, cb.gSpecMinMaterial));
            vmbOcclusion.w *= float(CompareMaterials(materialID, d11.z, cb.gSpecMinMaterial));

            fbits += vmbOcclusion.x * 16.0f;
            fbits += vmbOcclusion.y * 32.0f;
            fbits += vmbOcclusion.z * 64.0f;
            fbits += vmbOcclusion.w * 128.0f;

            vmbOcclusionWeights = Filtering::GetBilinearCustomWeights(vmbBilinearFilter, vmbOcclusion);
            vmbSpecAccumSpeed = Filtering::ApplyBilinearCustomWeights(d00.y, d10.y, d01.y, d11.y, vmbOcclusionWeights);

            float vmbFootprintQuality = Filtering::ApplyBilinearFilter(vmbOcclusion.x, vmbOcclusion.y, vmbOcclusion.z, vmbOcclusion.w, vmbBilinearFilter);
            vmbFootprintQuality = Math::Sqrt01(vmbFootprintQuality);

            float vmbSpecHistoryConfidence = vmbFootprintQuality;
            if (cb.gHasHistoryConfidence) {
                float confidence = saturate(t.gIn_SpecConfidence->sampleLinear(vmbPixelUv).x);
                vmbSpecHistoryConfidence = min(vmbSpecHistoryConfidence, confidence);
            }
            vmbSpecAccumSpeed *= lerp(vmbSpecHistoryConfidence, 1.0f, 1.0f / (1.0f + vmbSpecAccumSpeed));

            vmbAllowCatRom = dot(vmbOcclusion, float4(1.0f)) > 3.5f;
            vmbAllowCatRom = vmbAllowCatRom && smbAllowCatRom;
        }

        // How many radians can the virtual motion cover?
        float curvatureAngle, lobeHalfAngle;
        {
            float curvatureAngleTan = pixelSize * std::fabs(curvature);
            curvatureAngleTan *= max(vmbPixelsTraveled / max(NoV, 0.01f), 1.0f);
            curvatureAngleTan *= 2.0f;
            curvatureAngle = std::atan(curvatureAngleTan);

            float percentOfVolume = NRD_MAX_PERCENT_OF_LOBE_VOLUME / (1.0f + vmbSpecAccumSpeed);
            float lobeTanHalfAngle = ImportanceSampling::GetSpecularLobeTanHalfAngle(roughness, percentOfVolume);
            lobeTanHalfAngle = max(lobeTanHalfAngle, NRD_NORMAL_ENCODING_ERROR);
            hitDistanceToLobeSpreadInPixels *= lobeTanHalfAngle;
            lobeHalfAngle = std::atan(lobeTanHalfAngle);
        }

        // Virtual motion - confidence: parallax
        float parallaxWeight;
        {
            float hitDistForTrackingPrev = t.gPrev_SpecHitDistForTracking->sampleLinear(vmbPixelUv * cb.gResolutionScalePrev).x;
            float3 XvirtualPrev = GetXvirtual(hitDistForTrackingPrev, curvature, X, Xprev, N, V, roughness);

            float2 vmbPixelUvPrev = Geometry::GetScreenUv(cb.gWorldToClipPrev, XvirtualPrev);
            vmbPixelUvPrev = materialID == cb.gCameraAttachedReflectionMaterialID ? smbPixelUv : vmbPixelUvPrev;

            float r = min(hitDistForTracking, hitDistForTrackingPrev) * hitDistanceToLobeSpreadInPixels;
            r *= 0.5f;
            r = max(r, 0.1f * roughness);

            float d = length((vmbPixelUvPrev - vmbPixelUv) * cb.gRectSize);
            parallaxWeight = Math::LinearStep(r, 0.0f, d);
        }

        // Virtual motion - confidence: normal
        {
            float normalWeight = GetEncodingAwareNormalWeight(N, vmbN.xyz(), lobeHalfAngle, curvatureAngle, 0.0f);
            normalWeight = lerp(1.0f, normalWeight, Math::SmoothStep01(vmbPixelsTraveled));
            virtualHistoryConfidence *= normalWeight;
        }

        // Virtual motion - confidence: prev-prev tests (1 iteration)
        {
            float stepBetweenTaps = min(vmbPixelsTraveled * cb.gFramerateScale, 2.0f) + vmbPixelsTraveled / 1.0f;
            vmbDelta *= Math::Rsqrt(Math::LengthSquared(vmbDelta));
            vmbDelta /= cb.gRectSizePrev;

            float2 p = GetRelaxedRoughnessWeightParams(vmbN.w * vmbN.w, cb.gRoughnessFraction, REBLUR_ROUGHNESS_SENSITIVITY_IN_TA);
            for (int i = 1; i <= 1; i++) {
                float2 vmbPixelUvPrev = vmbPixelUv + vmbDelta * float(i) * stepBetweenTaps;
                float4 prevNR = NRD_FrontEnd_UnpackNormalAndRoughness(
                    t.gPrev_Normal_Roughness->sampleNearest(stochasticBilinear(rng, vmbPixelUvPrev, cb.gRectSizePrev) * cb.gResolutionScalePrev));

                float w = GetEncodingAwareNormalWeight(vmbN.xyz(), prevNR.xyz(), lobeHalfAngle, curvatureAngle * (1.0f + float(i) * stepBetweenTaps), 0.0f);
                w *= ComputeNonExponentialWeight(prevNR.w * prevNR.w, p.x, p.y);
                w = lerp(1.0f, w, saturate(stepBetweenTaps));
                w = IsInScreenNearest(vmbPixelUvPrev) != 0.0f ? w : 1.0f;
                virtualHistoryConfidence = min(virtualHistoryConfidence, w);
            }
        }

        virtualHistoryConfidence *= parallaxWeight;

        // Surface history confidence
        float surfaceHistoryConfidence;
        {
            float a = std::atan(smbParallaxInPixelsMax * pixelSize / length(X));
            float nonLinearAccumSpeed = 1.0f / (1.0f + smbSpecAccumSpeed);
            float hPrev = t.gHistory_Spec->sampleLinear(smbPixelUv * cb.gResolutionScalePrev).w;
            float h = lerp(hPrev, spec.w, nonLinearAccumSpeed) * hitDistNormalization;

            float tana0 = ImportanceSampling::GetSpecularLobeTanHalfAngle(roughnessModified, NRD_MAX_PERCENT_OF_LOBE_VOLUME);
            tana0 *= lerp(NoV, 1.0f, roughnessModified);
            tana0 *= nonLinearAccumSpeed;
            tana0 /= GetHitDistFactor(h, frustumSize) + NRD_EPS;

            float a0 = max(std::atan(tana0), NRD_NORMAL_ENCODING_ERROR);
            float f = Math::LinearStep(a0, 0.0f, a);
            surfaceHistoryConfidence = Math::Pow01(f, 4.0f);
            f = Math::LinearStep(0.8f, 0.9f, roughnessModified);
            surfaceHistoryConfidence = lerp(surfaceHistoryConfidence, 1.0f, f);
        }

        // Limit number of accumulated frames
        float smbSpecAccumSpeed_NoHistoryFix, vmbSpecAccumSpeed_NoHistoryFix;
        {
            float responsiveFactor = c.RemapRoughnessToResponsiveFactor(roughnessModified);
            float smc = GetSpecMagicCurve(roughnessModified);

            float2 f = float2(smbNoN, vmbNoN);
            float pw = lerp(32.0f, 1.0f, smc) * (1.0f - responsiveFactor);
            f = lerp(smc, 1.0f, responsiveFactor) * float2(Math::Pow01(f.x, pw), Math::Pow01(f.y, pw));

            float2 maxResponsiveFrameNum = float2(cb.gMaxAccumulatedFrameNum);
            maxResponsiveFrameNum *= f;
            maxResponsiveFrameNum = max(maxResponsiveFrameNum, float2((float)cb.gResponsiveAccumulationMinAccumulatedFrameNum));

            float2 maxFrameNum = cb.gMaxAccumulatedFrameNum * float2(surfaceHistoryConfidence, virtualHistoryConfidence);
            float2 maxFrameNum_NoHistoryFix = min(maxFrameNum, max(maxResponsiveFrameNum, float2(cb.gHistoryFixFrameNum)));

            smbSpecAccumSpeed_NoHistoryFix = min(smbSpecAccumSpeed, maxFrameNum_NoHistoryFix.x);
            vmbSpecAccumSpeed_NoHistoryFix = min(vmbSpecAccumSpeed, maxFrameNum_NoHistoryFix.y);

            maxFrameNum = min(maxFrameNum, maxResponsiveFrameNum);
            smbSpecAccumSpeed = min(smbSpecAccumSpeed, maxFrameNum.x);
            vmbSpecAccumSpeed = min(vmbSpecAccumSpeed, maxFrameNum.y);
        }

        // Virtual history amount
        {
            virtualHistoryAmount = 1.0f + (vmbSpecAccumSpeed - smbSpecAccumSpeed) / (1.0f + 0.5f * max(vmbSpecAccumSpeed, smbSpecAccumSpeed));
            virtualHistoryAmount = saturate(virtualHistoryAmount);
            if (!smbAllowCatRom || !vmbAllowCatRom) virtualHistoryAmount = step(0.5f, virtualHistoryAmount);
        }

        // Sample history
        float4 specHistory;
        float specFastHistory;
        {
            float2 uv = lerp(smbPixelUv, vmbPixelUv, virtualHistoryAmount);
            float4 occlusionWeights = lerp(smbOcclusionWeights, vmbOcclusionWeights, virtualHistoryAmount);
            bool allowCatRom = virtualHistoryAmount < 0.5f ? smbAllowCatRom : vmbAllowCatRom;
            HistoryFilter hf(saturate(uv) * cb.gRectSizePrev, cb.gResourceSizeInvPrev, occlusionWeights, allowCatRom);
            specHistory = hf.color(*t.gHistory_Spec);
            specFastHistory = hf.bilinear(*t.gHistory_SpecFast).x;
            specHistory = ReblurCtx::ClampNegativeToZero(specHistory);
            specFastHistory = max(specFastHistory, 0.0f);
        }

        // Accumulation
        specAccumSpeedCorrected = lerp(smbSpecAccumSpeed_NoHistoryFix, vmbSpecAccumSpeed_NoHistoryFix, virtualHistoryAmount);
        float specAccumSpeed = lerp(smbSpecAccumSpeed, vmbSpecAccumSpeed, virtualHistoryAmount);
        float specNonLinearAccumSpeed = 1.0f / (1.0f + specAccumSpeed);
        if (!specHasData) specNonLinearAccumSpeed *= lerp(1.0f - cb.gCheckerboardResolveAccumSpeed, 1.0f, specNonLinearAccumSpeed);

        float4 specResult = c.MixHistoryAndCurrent(specHistory, spec, specNonLinearAccumSpeed, roughness);

        // Firefly suppressor
        float specMaxRelativeIntensity = cb.gFireflySuppressorMinRelativeScale + REBLUR_FIREFLY_SUPPRESSOR_MAX_RELATIVE_INTENSITY / (specAccumSpeed + 1.0f);
        float specAntifireflyFactor = specAccumSpeed * cb.gMaxBlurRadius * REBLUR_FIREFLY_SUPPRESSOR_RADIUS_SCALE;
        specAntifireflyFactor /= 1.0f + specAntifireflyFactor;
        {
            float specLumaResult = ReblurCtx::GetLuma(specResult);
            float specLumaClamped = min(specLumaResult, ReblurCtx::GetLuma(specHistory) * specMaxRelativeIntensity);
            specLumaClamped = lerp(specLumaResult, specLumaClamped, specAntifireflyFactor);
            specResult = ReblurCtx::ChangeLuma(specResult, specLumaClamped);

            float specHitDistMaxRelativeIntensity = 1.2f + 1.0f / (specAccumSpeed + 1.0f);
            specResult.w = lerp(specResult.w, min(specResult.w, specHistory.w * specHitDistMaxRelativeIntensity), specAntifireflyFactor);
        }
        t.gOut_Spec->store(px, py, specResult);

        {  // Fast history
            float maxFastAccumulatedFrameNum = cb.gMaxFastAccumulatedFrameNum;
            if (materialID == cb.gStrandMaterialID) maxFastAccumulatedFrameNum = max(maxFastAccumulatedFrameNum, cb.gMaxAccumulatedFrameNum / 5);

            float specHistoryConfidence = lerp(surfaceHistoryConfidence, virtualHistoryConfidence, virtualHistoryAmount);
            float specFastNonLinearAccumSpeed = c.GetNonLinearAccumSpeed(specAccumSpeed, maxFastAccumulatedFrameNum, specHistoryConfidence, specHasData);
            float specFastResult = lerp(specFastHistory, ReblurCtx::GetLuma(spec), specFastNonLinearAccumSpeed);

            float specFastClamped = min(specFastResult, ReblurCtx::GetLuma(specHistory) * specMaxRelativeIntensity * REBLUR_FIREFLY_SUPPRESSOR_FAST_RELATIVE_INTENSITY);
            specFastResult = lerp(specFastResult, specFastClamped, specAntifireflyFactor);
            t.gOut_SpecFast->store(px, py, float4(specFastResult));
        }
    }

    t.gOut_Data2->storeUint(px, py, ReblurCtx::PackData2(fbits, curvature, virtualHistoryAmount, smbAllowCatRom, c.signal));

    // ---------------------------------------------------------------------------------------------- Diffuse
    if (!c.hasDiff()) diffAccumSpeed = 0.0f;  // TA:972-974
    else {
        float diffHistoryConfidence = smbFootprintQuality;
        if (cb.gHasHistoryConfidence) {
            float confidence = saturate(t.gIn_DiffConfidence->sampleLinear(smbPixelUv).x);
            diffHistoryConfidence = min(diffHistoryConfidence, confidence);
        }
        diffAccumSpeed *= lerp(diffHistoryConfidence, 1.0f, 1.0f / (1.0f + diffAccumSpeed));

        const bool diffHasData = cb.gDiffCheckerboard == 2 || Sequence::CheckerBoard((uint32_t)px, (uint32_t)py, cb.gFrameIndex) == cb.gDiffCheckerboard;
        float4 diff = t.gIn_Diff->load(px, py);

        float4 diffHistory;
        float diffFastHistory;
        {
            HistoryFilter hf(saturate(smbPixelUv) * cb.gRectSizePrev, cb.gResourceSizeInvPrev, smbOcclusionWeights, smbAllowCatRom);
            diffHistory = hf.color(*t.gHistory_Diff);
            diffFastHistory = hf.bilinear(*t.gHistory_DiffFast).x;
            diffHistory = ReblurCtx::ClampNegativeToZero(diffHistory);
            diffFastHistory = max(diffFastHistory, 0.0f);
        }

        float diffNonLinearAccumSpeed = 1.0f / (1.0f + diffAccumSpeed);
        if (!diffHasData) diffNonLinearAccumSpeed *= lerp(1.0f - cb.gCheckerboardResolveAccumSpeed, 1.0f, diffNonLinearAccumSpeed);
        float4 diffResult = c.MixHistoryAndCurrent(diffHistory, diff, diffNonLinearAccumSpeed);

        float diffMaxRelativeIntensity = cb.gFireflySuppressorMinRelativeScale + REBLUR_FIREFLY_SUPPRESSOR_MAX_RELATIVE_INTENSITY / (diffAccumSpeed + 1.0f);
        float diffAntifireflyFactor = diffAccumSpeed * cb.gMaxBlurRadius * REBLUR_FIREFLY_SUPPRESSOR_RADIUS_SCALE;
        diffAntifireflyFactor /= 1.0f + diffAntifireflyFactor;

        float diffLumaResult = ReblurCtx::GetLuma(diffResult);
        float diffLumaClamped = min(diffLumaResult, ReblurCtx::GetLuma(diffHistory) * diffMaxRelativeIntensity);
        diffLumaClamped = lerp(diffLumaResult, diffLumaClamped, diffAntifireflyFactor);
        diffResult = ReblurCtx::ChangeLuma(diffResult, diffLumaClamped);

        float diffHitDistMaxRelativeIntensity = 1.2f + 1.0f / (diffAccumSpeed + 1.0f);
        diffResult.w = lerp(diffResult.w, min(diffResult.w, diffHistory.w * diffHitDistMaxRelativeIntensity), diffAntifireflyFactor);

        t.gOut_Diff->store(px, py, diffResult);

        {  // Fast history
            float diffFastAccumSpeed = min(diffAccumSpeed, cb.gMaxFastAccumulatedFrameNum);
            float diffFastNonLinearAccumSpeed = 1.0f / (1.0f + diffFastAccumSpeed);
            if (!diffHasData) diffFastNonLinearAccumSpeed *= lerp(1.0f - cb.gCheckerboardResolveAccumSpeed, 1.0f, diffFastNonLinearAccumSpeed);
            float diffFastResult = lerp(diffFastHistory, ReblurCtx::GetLuma(diff), diffFastNonLinearAccumSpeed);
            float diffFastClamped = min(diffFastResult, ReblurCtx::GetLuma(diffHistory) * diffMaxRelativeIntensity * REBLUR_FIREFLY_SUPPRESSOR_FAST_RELATIVE_INTENSITY);
            diffFastResult = lerp(diffFastResult, diffFastClamped, diffAntifireflyFactor);
            t.gOut_DiffFast->store(px, py, float4(diffFastResult));
        }
    }

    float2 d1 = ReblurCtx::PackData1(diffAccumSpeed, specAccumSpeedCorrected, c.signal);
    t.gOut_Data1->store(px, py, float4(d1.x, d1.y, 0, 0));
}

void reblurTemporalAccumulation(const ReblurCB& cb, const TaTextures& t, int gridW, int gridH, int signal) {
    ReblurCtx c(cb, signal);
    const int W = gridW * 8, H = gridH * 16;
#pragma omp parallel for schedule(dynamic, 4)
    for (int py = 0; py < H; py++)
        for (int px = 0; px < W; px++) {
            float isSky = t.gIn_Tiles->load(px >> 4, py >> 4).x;
            if (isSky != 0.0f || px > cb.gRectSizeMinusOne.x || py > cb.gRectSizeMinusOne.y) continue;
            taPixel(c, t, px, py);
        }
}

// ------------------------------------------------------------------------------------------------------------
// History fix
// ------------------------------------------------------------------------------------------------------------
namespace {

float hfSmemLuma(const ReblurCtx& c, const HfTextures& t, const Tex& fast, int gx, int gy) {
    gx = clampi(gx, 0, c.cb.gRectSizeMinusOne.x);
    gy = clampi(gy, 0, c.cb.gRectSizeMinusOne.y);
    float viewZ = c.UnpackViewZ(t.gIn_ViewZ->load(gx, gy).x);
    return !c.IsInDenoisingRange(viewZ) ? REBLUR_INVALID : fast.load(gx, gy).x;
}

// value each lane contributes to the quad exchange: ( frameNum < gHistoryFixFrameNum )
float2 hfStridePreQuad(const ReblurCtx& c, const HfTextures& t, int px, int py, float2* frameNumOut, float* viewZOut) {
    float4 d = t.gIn_Data1->load(px, py);
    float2 frameNum = ReblurCtx::UnpackData1(float2(d.x, d.y), c.signal);
    float viewZ = c.UnpackViewZ(t.gIn_ViewZ->load(px, py).x);
    if (!c.IsInDenoisingRange(viewZ)) frameNum = float2((float)REBLUR_MAX_ACCUM_FRAME_NUM);
    if (frameNumOut) *frameNumOut = frameNum;
    if (viewZOut) *viewZOut = viewZ;
    return float2(float(frameNum.x < c.cb.gHistoryFixFrameNum), float(frameNum.y < c.cb.gHistoryFixFrameNum));
}

}  // namespace

template <int LOBE>
static void hfLobe(const ReblurCtx& c, const HfTextures& t, int px, int py, float strideIn, float frameNum, float frameNumAvgNorm, float viewZ, float materialID,
                   float3 N, float roughness, float3 Nv, float3 Xv, float frustumSize, float2 pixelUv) {
    const ReblurCB& cb = c.cb;
    const Tex& IN = LOBE == DIFF ? *t.gIn_Diff : *t.gIn_Spec;
    const Tex& FAST = LOBE == DIFF ? *t.gIn_DiffFast : *t.gIn_SpecFast;
    Tex& OUT = LOBE == DIFF ? *t.gOut_Diff : *t.gOut_Spec;
    Tex& OUT_FAST = LOBE == DIFF ? *t.gOut_DiffFast : *t.gOut_SpecFast;
    const float MIN_MATERIAL = LOBE == DIFF ? cb.gDiffMinMaterial : cb.gSpecMinMaterial;
    const int BORDER = 4;  // REBLUR_ANTI_FIREFLY_FILTER_RADIUS (NRD_SUPPORTS_ANTIFIREFLY = 1)

    float4 v = IN.load(px, py);
    float smc = LOBE == DIFF ? 1.0f : GetSpecMagicCurve(roughness);
    float nonLinearAccumSpeed = 1.0f / (1.0f + frameNum);

    float hitDistScale = _REBLUR_GetHitDistanceNormalization(viewZ, cb.gHitDistSettings.xyz(), LOBE == DIFF ? 1.0f : roughness);
    float hitDist = v.w * hitDistScale;
    if (LOBE == SPEC) hitDist = lerp(t.gIn_SpecHitDistForTracking->load(px, py).x, hitDist, smc);
    float hitDistFactor = GetHitDistFactor(hitDist, frustumSize);
    hitDist = LOBE == DIFF ? v.w : saturate(hitDist / hitDistScale);

    float stride = strideIn;
    stride *= lerp(0.25f + 0.75f * Math::Sqrt01(hitDistFactor), 1.0f, nonLinearAccumSpeed);
    if (LOBE == SPEC) stride *= lerp(0.25f, 1.0f, smc);
    stride = hlsl_round(stride);

    if (stride != 0.0f) {
        float normalWeightParam = GetNormalWeightParam(nonLinearAccumSpeed, cb.gLobeAngleFraction, LOBE == DIFF ? 1.0f : roughness);
        float2 geometryWeightParams = GetGeometryWeightParams(cb.gPlaneDistSensitivity, frustumSize, Xv, Nv);
        float2 hitDistanceWeightParams = GetHitDistanceWeightParams(hitDist, nonLinearAccumSpeed);
        float2 relaxedRoughnessWeightParams = GetRelaxedRoughnessWeightParams(roughness * roughness, std::sqrt(cb.gRoughnessFraction));

        float sum = 1.0f + frameNum;
        v *= sum;

        for (int j = -2; j <= 2; j++)
            for (int i = -2; i <= 2; i++) {
                if (i == 0 && j == 0) continue;
                if (std::abs(i) + std::abs(j) == 4) continue;

                float2 uv = pixelUv + float2((float)i, (float)j) * stride * cb.gRectSizeInv;
                uv = MirrorUv(uv);
                float2 posf = uv * cb.gRectSize;
                int2 pos = int2((int)posf.x, (int)posf.y);

                float zs = c.UnpackViewZ(t.gIn_ViewZ->load(pos).x);
                float3 Xvs = Geometry::ReconstructViewPosition(uv, cb.gFrustum, zs, cb.gOrthoMode);

                float materialIDs;
                float4 Ns = unpackNR(*t.gIn_Normal_Roughness, pos.x, pos.y, materialIDs);

                float angle = Math::AcosApproxPositive(dot(Ns.xyz(), N));
                float NoX = dot(Nv, Xvs);

                float w = float(CompareMaterials(materialID, materialIDs, MIN_MATERIAL));
                w *= ComputeExponentialWeight(angle, normalWeightParam, 0.0f);
                if (LOBE == SPEC) w *= ComputeExponentialWeight(Ns.w * Ns.w, relaxedRoughnessWeightParams.x, relaxedRoughnessWeightParams.y);

                float4 d1 = t.gIn_Data1->load(pos);
                float2 fn = ReblurCtx::UnpackData1(float2(d1.x, d1.y), c.signal);
                w *= 1.0f + (LOBE == DIFF ? fn.x : fn.y);

                w = c.ApplyGeometryWeightLast(w, zs, NoX, geometryWeightParams);

                float4 smp = IN.load(pos);
                smp = w == 0.0f ? float4(0.0f) : smp;

                w *= ComputeExponentialWeight(smp.w, hitDistanceWeightParams.x, hitDistanceWeightParams.y);

                sum += w;
                v += smp * w;
            }

        sum = Math::PositiveRcp(sum);
        v *= sum;
    }

    float luma = ReblurCtx::GetLuma(v);

    float f = frameNumAvgNorm;
    if (LOBE == SPEC) f = lerp(1.0f, f, smc);

    float fastCenter = hfSmemLuma(c, t, FAST, px, py);
    fastCenter = lerp(luma, fastCenter, f);
    OUT_FAST.store(px, py, float4(fastCenter));

    // Local variance
    float fastM1 = fastCenter, fastM2 = fastCenter * fastCenter;
    float antiFireflyM1 = 0.0f, antiFireflyM2 = 0.0f;
    for (int j = -BORDER; j <= BORDER; j++)
        for (int i = -BORDER; i <= BORDER; i++) {
            if (i == 0 && j == 0) continue;
            float d = hfSmemLuma(c, t, FAST, px + i, py + j);
            d = d == REBLUR_INVALID ? fastCenter : d;
            if (std::abs(i) <= 2 && std::abs(j) <= 2) {
                fastM1 += d;
                fastM2 += d * d;
            }
            if (!(std::abs(i) <= 1 && std::abs(j) <= 1)) {
                antiFireflyM1 += d;
                antiFireflyM2 += d * d;
            }
        }

    if (cb.gAntiFirefly != 0.0f) {
        float invNorm = 1.0f / ((BORDER * 2 + 1) * (BORDER * 2 + 1) - 3 * 3);
        antiFireflyM1 *= invNorm;
        antiFireflyM2 *= invNorm;
        float sigma = GetStdDev(antiFireflyM1, antiFireflyM2) * REBLUR_ANTI_FIREFLY_SIGMA_SCALE;
        luma = clamp(luma, antiFireflyM1 - sigma, antiFireflyM1 + sigma);
    }

    {
        float invNorm = 1.0f / 25.0f;
        fastM1 *= invNorm;
        fastM2 *= invNorm;
        float scale = cb.gFastHistoryClampingSigmaScale;
        if (LOBE == SPEC && materialID == cb.gStrandMaterialID) scale = max(scale, 3.0f);
        float sigma = GetStdDev(fastM1, fastM2) * scale;
        float lumaClamped = clamp(luma, fastM1 - sigma, fastM1 + sigma);
        luma = lerp(lumaClamped, luma, 1.0f / (1.0f + float(cb.gMaxFastAccumulatedFrameNum < cb.gMaxAccumulatedFrameNum) * frameNum * 2.0f));
    }

    v = ReblurCtx::ChangeLuma(v, luma);
    OUT.store(px, py, v);
}

void reblurHistoryFix(const ReblurCB& cb, const HfTextures& t, int gridW, int gridH, bool quads, int signal) {
    ReblurCtx c(cb, signal);
    const int W = gridW * 8, H = gridH * 16;
#pragma omp parallel for schedule(dynamic, 4)
    for (int py = 0; py < H; py++)
        for (int px = 0; px < W; px++) {
            float isSky = t.gIn_Tiles->load(px >> 4, py >> 4).x;
            if (isSky != 0.0f) continue;

            float2 frameNum;
            float viewZ;
            float2 stride = hfStridePreQuad(c, t, px, py, &frameNum, &viewZ);
            if (quads) {
                float2 d10 = hfStridePreQuad(c, t, px ^ 1, py, nullptr, nullptr);
                float2 d01 = hfStridePreQuad(c, t, px ^ 2, py, nullptr, nullptr);
                float2 avg = (d10 + d01 + stride) / 3.0f;
                stride = min(stride, avg);
            }
            if (!c.IsInDenoisingRange(viewZ) || px > cb.gRectSizeMinusOne.x || py > cb.gRectSizeMinusOne.y) continue;

            float materialID;
            float4 nr = unpackNR(*t.gIn_Normal_Roughness, px, py, materialID);
            float3 N = nr.xyz();
            float roughness = nr.w;

            float frustumSize = GetFrustumSize(cb.gMinRectDimMulUnproject, cb.gOrthoMode, viewZ);
            float2 pixelUv = float2(px + 0.5f, py + 0.5f) * cb.gRectSizeInv;
            float3 Xv = Geometry::ReconstructViewPosition(pixelUv, cb.gFrustum, viewZ, cb.gOrthoMode);
            float3 Nv = Geometry::RotateVectorInverse(cb.gViewToWorld, N);

            float invHistoryFixFrameNum = 1.0f / max(cb.gHistoryFixFrameNum, NRD_EPS);
            float2 frameNumAvgNorm = saturate(frameNum * invHistoryFixFrameNum);

            stride /= 1.0f + 1.0f;
            stride *= 2.0f / 2.0f;  // REBLUR_HISTORY_FIX_FILTER_RADIUS = 2
            stride *= materialID == cb.gHistoryFixAlternatePixelStrideMaterialID ? cb.gHistoryFixAlternatePixelStride : cb.gHistoryFixBasePixelStride;

            if (c.hasDiff()) hfLobe<DIFF>(c, t, px, py, stride.x, frameNum.x, frameNumAvgNorm.x, viewZ, materialID, N, roughness, Nv, Xv, frustumSize, pixelUv);
            if (c.hasSpec()) hfLobe<SPEC>(c, t, px, py, stride.y, frameNum.y, frameNumAvgNorm.y, viewZ, materialID, N, roughness, Nv, Xv, frustumSize, pixelUv);
        }
}

// ------------------------------------------------------------------------------------------------------------
// Temporal stabilization
// ------------------------------------------------------------------------------------------------------------
namespace {
float tsSmemLuma(const ReblurCtx& c, const TsTextures& t, const Tex& in, int gx, int gy) {
    gx = clampi(gx, 0, c.cb.gRectSizeMinusOne.x);
    gy = clampi(gy, 0, c.cb.gRectSizeMinusOne.y);
    float viewZ = c.UnpackViewZ(t.gIn_ViewZ->load(gx, gy).x);
    return !c.IsInDenoisingRange(viewZ) ? REBLUR_INVALID : ReblurCtx::GetLuma(in.load(gx, gy));
}
void tsMoments(const ReblurCtx& c, const TsTextures& t, const Tex& in, int px, int py, float& luma, float& m1, float& sigma) {
    luma = tsSmemLuma(c, t, in, px, py);
    m1 = luma;
    float m2 = luma * luma;
    for (int j = 0; j <= 2; j++)
        for (int i = 0; i <= 2; i++) {
            if (i == 1 && j == 1) continue;
            float d = tsSmemLuma(c, t, in, px - 1 + i, py - 1 + j);
            d = d == REBLUR_INVALID ? luma : d;
            m1 += d;
            m2 += d * d;
        }
    m1 /= 9.0f;
    m2 /= 9.0f;
    sigma = GetStdDev(m1, m2);
}
}  // namespace

static void tsPixel(const ReblurCtx& c, const TsTextures& t, int px, int py) {
    const ReblurCB& cb = c.cb;
    const float3 cameraDelta = cb.gCameraDelta.xyz();

    float viewZ = c.UnpackViewZ(t.gIn_ViewZ->load(px, py).x);
    if (!c.IsInDenoisingRange(viewZ)) return;

    float2 pixelUv = float2(px + 0.5f, py + 0.5f) * cb.gRectSizeInv;
    float3 Xv = Geometry::ReconstructViewPosition(pixelUv, cb.gFrustum, viewZ, cb.gOrthoMode);
    float3 X = Geometry::RotateVector(cb.gViewToWorld, Xv);

    float4 inMv = t.gInOut_Mv->load(px, py);
    float3 mv = inMv.xyz() * cb.gMvScale.xyz();
    float3 Xprev = X;
    float2 smbPixelUv = pixelUv + mv.xy();
    if (cb.gMvScale.w == 0.0f) {
        if (cb.gMvScale.z == 0.0f) mv.z = Geometry::AffineTransform(cb.gWorldToViewPrev, X).z - viewZ;
        float viewZprev = viewZ + mv.z;
        float3 Xvprevlocal = Geometry::ReconstructViewPosition(smbPixelUv, cb.gFrustumPrev, viewZprev, cb.gOrthoMode);
        Xprev = Geometry::RotateVectorInverse(cb.gWorldToViewPrev, Xvprevlocal) + cameraDelta;
    } else {
        Xprev += mv;
        smbPixelUv = Geometry::GetScreenUv(cb.gWorldToClipPrev, Xprev);
    }

    float materialID;
    float4 nr = unpackNR(*t.gIn_Normal_Roughness, px, py, materialID);
    float3 N = nr.xyz();
    float roughness = nr.w;

    uint32_t bits;
    bool smbAllowCatRom;
    float4 d1 = t.gIn_Data1->load(px, py);
    float2 data1 = ReblurCtx::UnpackData1(float2(d1.x, d1.y), c.signal);
    float2 data2 = ReblurCtx::UnpackData2(t.gIn_Data2->loadUint(px, py), bits, smbAllowCatRom, c.signal);

    Filtering::Bilinear smbBilinearFilter = Filtering::GetBilinearFilter(smbPixelUv, cb.gRectSizePrev);
    float4 smbOcclusion = float4(float((bits & 1u) != 0), float((bits & 2u) != 0), float((bits & 4u) != 0), float((bits & 8u) != 0));
    float4 smbOcclusionWeights = Filtering::GetBilinearCustomWeights(smbBilinearFilter, smbOcclusion);
    float smbFootprintQuality = Filtering::ApplyBilinearFilter(smbOcclusion.x, smbOcclusion.y, smbOcclusion.z, smbOcclusion.w, smbBilinearFilter);
    smbFootprintQuality = Math::Sqrt01(smbFootprintQuality);

    // Diffuse
    if (c.hasDiff()) {
        float diffLuma, diffLumaM1, diffLumaSigma;
        tsMoments(c, t, *t.gIn_Diff, px, py, diffLuma, diffLumaM1, diffLumaSigma);

        if (data1.x < cb.gHistoryFixFrameNum) diffLuma = min(diffLuma, diffLumaM1 * (1.2f + 1.0f / (1.0f + data1.x)));

        HistoryFilter hf(saturate(smbPixelUv) * cb.gRectSizePrev, cb.gResourceSizeInvPrev, smbOcclusionWeights, smbAllowCatRom);
        float diffLumaHistory = hf.color(*t.gHistory_DiffLumaStabilized).x;
        diffLumaHistory = max(diffLumaHistory, 0.0f);

        float diffAntilag = c.ComputeAntilag(diffLumaHistory, diffLumaM1, diffLumaSigma, smbFootprintQuality * data1.x);
        float diffMinAccumSpeed = min(data1.x, cb.gHistoryFixFrameNum) * 1.0f;
        data1.x = lerp(diffMinAccumSpeed, data1.x, diffAntilag);

        float2 params = c.GetTemporalAccumulationParams(smbFootprintQuality, data1.x, diffAntilag);
        float diffHistoryWeight = params.x;
        diffHistoryWeight *= float(pixelUv.x >= cb.gSplitScreen);
        diffHistoryWeight *= float(smbPixelUv.x >= cb.gSplitScreenPrev);

        diffLumaHistory = Color::Clamp(diffLumaM1, diffLumaSigma * params.y, diffLumaHistory);
        float diffLumaStabilized = lerp(diffLuma, diffLumaHistory, min(diffHistoryWeight, cb.gStabilizationStrength));

        float4 diff = t.gIn_Diff->load(px, py);
        diff = ReblurCtx::ChangeLuma(diff, diffLumaStabilized);
        diff.w = cb.gReturnHistoryLengthInsteadOfOcclusion ? data1.x : diff.w;

        t.gOut_Diff->store(px, py, diff);
        t.gOut_DiffLumaStabilized->store(px, py, float4(diffLumaStabilized));
    }

    // Specular
    if (c.hasSpec()) {
        float specLuma, specLumaM1, specLumaSigma;
        tsMoments(c, t, *t.gIn_Spec, px, py, specLuma, specLumaM1, specLumaSigma);

        if (data1.y < cb.gHistoryFixFrameNum) specLuma = min(specLuma, specLumaM1 * (1.2f + 1.0f / (1.0f + data1.y)));

        float hitDistForTracking = t.gIn_SpecHitDistForTracking->load(px, py).x;
        float virtualHistoryAmount = data2.x;
        float curvature = data2.y;

        float3 V = c.GetViewVector(X);
        float3 Xvirtual = GetXvirtual(hitDistForTracking, curvature, X, Xprev, N, V, roughness);
        float2 vmbPixelUv = Geometry::GetScreenUv(cb.gWorldToClipPrev, Xvirtual);
        vmbPixelUv = materialID == cb.gCameraAttachedReflectionMaterialID ? pixelUv : vmbPixelUv;

        Filtering::Bilinear vmbBilinearFilter = Filtering::GetBilinearFilter(vmbPixelUv, cb.gRectSizePrev);
        float4 vmbOcclusion = float4(float((bits & 16u) != 0), float((bits & 32u) != 0), float((bits & 64u) != 0), float((bits & 128u) != 0));
        float4 vmbOcclusionWeights = Filtering::GetBilinearCustomWeights(vmbBilinearFilter, vmbOcclusion);

        bool vmbAllowCatRom = dot(vmbOcclusion, float4(1.0f)) > 3.5f;
        vmbAllowCatRom = vmbAllowCatRom && smbAllowCatRom;

        float vmbFootprintQuality = Filtering::ApplyBilinearFilter(vmbOcclusion.x, vmbOcclusion.y, vmbOcclusion.z, vmbOcclusion.w, vmbBilinearFilter);
        vmbFootprintQuality = Math::Sqrt01(vmbFootprintQuality);

        float2 uv = lerp(smbPixelUv, vmbPixelUv, virtualHistoryAmount);
        float4 occlusionWeights = lerp(smbOcclusionWeights, vmbOcclusionWeights, virtualHistoryAmount);
        bool allowCatRom = virtualHistoryAmount < 0.5f ? smbAllowCatRom : vmbAllowCatRom;

        HistoryFilter hf(saturate(uv) * cb.gRectSizePrev, cb.gResourceSizeInvPrev, occlusionWeights, allowCatRom);
        float specLumaHistory = hf.color(*t.gHistory_SpecLumaStabilized).x;
        specLumaHistory = max(specLumaHistory, 0.0f);

        float footprintQuality = lerp(smbFootprintQuality, vmbFootprintQuality, virtualHistoryAmount);
        float specAntilag = c.ComputeAntilag(specLumaHistory, specLumaM1, specLumaSigma, footprintQuality * data1.y);
        float specMinAccumSpeed = min(data1.y, cb.gHistoryFixFrameNum) * 1.0f;
        data1.y = lerp(specMinAccumSpeed, data1.y, specAntilag);

        float2 params = c.GetTemporalAccumulationParams(footprintQuality, data1.y, specAntilag);
        float specHistoryWeight = params.x;
        specHistoryWeight *= float(pixelUv.x >= cb.gSplitScreen);
        specHistoryWeight *= virtualHistoryAmount != 1.0f ? float(smbPixelUv.x >= cb.gSplitScreenPrev) : 1.0f;
        specHistoryWeight *= virtualHistoryAmount != 0.0f ? float(vmbPixelUv.x >= cb.gSplitScreenPrev) : 1.0f;

        float responsiveFactor = c.RemapRoughnessToResponsiveFactor(roughness);
        float smc = GetSpecMagicCurve(roughness);
        float acceleration = lerp(smc, 1.0f, 0.5f + responsiveFactor * 0.5f);
        if (materialID == cb.gStrandMaterialID) acceleration = min(acceleration, 0.5f);
        specHistoryWeight *= acceleration;

        specLumaHistory = Color::Clamp(specLumaM1, specLumaSigma * params.y, specLumaHistory);
        float specLumaStabilized = lerp(specLuma, specLumaHistory, min(specHistoryWeight, cb.gStabilizationStrength));

        float4 spec = t.gIn_Spec->load(px, py);
        spec = ReblurCtx::ChangeLuma(spec, specLumaStabilized);
        spec.w = cb.gReturnHistoryLengthInsteadOfOcclusion ? data1.y : spec.w;

        t.gOut_Spec->store(px, py, spec);
        t.gOut_SpecLumaStabilized->store(px, py, float4(specLumaStabilized));
    }

    t.gOut_InternalData->storeUint(px, py, c.PackInternalData(data1.x, data1.y, materialID));
}

void reblurTemporalStabilization(const ReblurCB& cb, const TsTextures& t, int gridW, int gridH, int signal) {
    ReblurCtx c(cb, signal);
    const int W = gridW * 8, H = gridH * 16;
#pragma omp parallel for schedule(dynamic, 4)
    for (int py = 0; py < H; py++)
        for (int px = 0; px < W; px++) {
            float isSky = t.gIn_Tiles->load(px >> 4, py >> 4).x;
            if (isSky != 0.0f || px > cb.gRectSizeMinusOne.x || py > cb.gRectSizeMinusOne.y) continue;
            tsPixel(c, t, px, py);
        }
}

// ------------------------------------------------------------------------------------------------------------
void clearTexture(Tex& out) {
#pragma omp parallel for schedule(static)
    for (int y = 0; y < out.h; y++) memset(out.data + (size_t)y * out.pitch, 0, (size_t)out.w * out.bytesPerTexel());
}

}  // namespace orc
