// REBLUR spatial passes on sm_100a: ClassifyTiles, PrePass, Blur, PostBlur (NRD_SIGNAL=BOTH, NRD_MODE=RADIANCE).
//
// What they replace: External/NRD/Shaders/REBLUR_ClassifyTiles.cs.hlsl:21-55, REBLUR_PrePass.cs.hlsl:21-86,
// REBLUR_Blur.cs.hlsl:21-92, REBLUR_PostBlur.cs.hlsl:21-95 and their shared body
// REBLUR_Common_SpatialFilter.hlsli:59-336 (8-tap rotated Poisson gather with material / normal / roughness /
// plane-distance / hit-distance edge stopping; diffuse taps in screen space, specular taps on a lobe-aligned
// world-space basis projected through gViewToClip).
//
// Mapping to the GPU: one thread per pixel, 32x8 CTAs so that a warp owns 32 consecutive pixels of one row
// (centre-pixel traffic is one 128/256-byte line per plane per warp; the reference's SM6.0 quads — 4 consecutive
// lanes of a row — are `__shfl_xor 1 / 2`). The 16 taps per pixel are data-dependent gathers within a <= 60 px
// radius: they hit L1/L2, not HBM, so no shared-memory staging is attempted (a TMA box cannot follow a per-pixel
// rotated, mirrored, lobe-projected footprint). Diffuse and specular share the centre set-up and run back to back
// in one kernel so G-buffer lines fetched by one lobe are still in L1 for the other.
#include "reblur_common.cuh"

namespace nrdk {

namespace {

enum { PRE_PASS = 0, BLUR = 1, POST_BLUR = 2 };
enum { DIFF = 0, SPEC = 1 };

constexpr int BLOCK_W = 32, BLOCK_H = 8;

// g_Special8 (Common.hlsli:207-218) as compile-time immediates: the tap loop is fully unrolled, so offsets fold into
// FFMA immediates and the per-tap gaussian weight exp(-0.66 z^2), z in {1, 0.5}, becomes a literal.
constexpr float kQ = 0.35355339059327373f;  // 0.25 * sqrt(2)
__device__ constexpr float kSpecial8X[8] = {-1.0f, 0.0f, 1.0f, 0.0f, -kQ, kQ, kQ, -kQ};
__device__ constexpr float kSpecial8Y[8] = {0.0f, 1.0f, 0.0f, -1.0f, kQ, kQ, -kQ, -kQ};
constexpr float kGaussOuter = 0.5168513209367337f;  // exp(-0.66f * 1.0 * 1.0)
constexpr float kGaussInner = 0.8478936985286916f;  // exp(-0.66f * 0.5 * 0.5)

struct Center {
    int px, py;
    float viewZ, materialID, roughness, NoV, frustumSize;
    float3 N, Nv, Xv, Vv;
    float2 pixelUv, nonLinearAccumSpeed, data1;
    float4 rotator;
};

NRD_DEV void setupCenter(const ReblurConstants& cb, Center& s, const TexNR& normalRoughness, const float* baseRotator) {
    float4 nr = unpackNormalRoughness(normalRoughness.loadRaw(s.px, s.py), s.materialID);
    s.N = xyz(nr);
    s.Nv = rotateInverse(cb.viewToWorld, s.N);
    s.roughness = nr.w;
    s.pixelUv = make_float2(s.px + 0.5f, s.py + 0.5f) * make_float2(cb.rectSizeInv[0], cb.rectSizeInv[1]);
    s.Xv = reconstructViewPosition(s.pixelUv, cb.frustum, s.viewZ, cb.orthoMode);
    s.Vv = viewVector(cb, s.Xv, true);
    s.NoV = fabsf(dot(s.Nv, s.Vv));
    s.frustumSize = frustumSizeAt(cb.minRectDimMulUnproject, cb.orthoMode, s.viewZ);
    s.rotator = make_float4(baseRotator[0], baseRotator[1], baseRotator[2], baseRotator[3]);
}

// One lobe of one spatial pass for one pixel
template <int PASS, int LOBE>
NRD_DEV void spatialFilter(const ReblurConstants& cb, const Center& s, const TexR32F& viewZTex, const TexNR& nrTex, const TexRGBA16F& input,
                           const TexRGBA16F& output, const TexR16F* outSpecHitDistForTracking, const TexRGBA16F* outputCopy, bool temporalStabilization, bool robustMirrorTest) {
    const float ROUGHNESS = LOBE == DIFF ? 1.0f : s.roughness;
    const float NLAS = LOBE == DIFF ? s.nonLinearAccumSpeed.x : s.nonLinearAccumSpeed.y;
    const float MIN_MATERIAL = LOBE == DIFF ? cb.diffMinMaterial : cb.specMinMaterial;
    const float MAX_BLUR_RADIUS = PASS == PRE_PASS ? (LOBE == DIFF ? cb.diffPrepassBlurRadius : cb.specPrepassBlurRadius) : cb.maxBlurRadius;
    constexpr bool SCREEN_SPACE = PASS == PRE_PASS || LOBE == DIFF;
    const float2 rectSize = make_float2(cb.rectSize[0], cb.rectSize[1]);
    const float2 rectSizeInv = make_float2(cb.rectSizeInv[0], cb.rectSizeInv[1]);

    float sum = 1.0f;
    float4 result = input.load(s.px, s.py);

    if (PASS != PRE_PASS || MAX_BLUR_RADIUS != 0.0f) {
        Rng rng;
        if (PASS == PRE_PASS && LOBE == SPEC) rng.init((uint32_t)s.px, (uint32_t)s.py, cb.frameIndex);

        constexpr float radiusScale = PASS == POST_BLUR ? 2.0f : 1.0f;
        constexpr float fractionScale = PASS == PRE_PASS ? 2.0f : (PASS == BLUR ? 1.0f : 0.5f);

        float4 Dv = specularDominantDirectionG2(s.Nv, s.Vv, ROUGHNESS);
        float NoD = fabsf(dot(s.Nv, xyz(Dv)));
        float smc = specMagicCurve(ROUGHNESS, 0.5f);

        float hitDistScale = hitDistanceNormalization(s.viewZ, cb.hitDistSettings, ROUGHNESS);
        float hitDist = result.w * hitDistScale;
        float hdFactor = hitDistFactor(hitDist, s.frustumSize);

        float areaFactor = PASS == PRE_PASS ? hdFactor : hdFactor * NLAS;
        float blurRadius = radiusScale * sqrt01(areaFactor);
        blurRadius = saturate(blurRadius) * MAX_BLUR_RADIUS * smc;
        blurRadius = fmaxf(blurRadius, cb.minBlurRadius * smc);

        if (PASS == PRE_PASS && LOBE == SPEC) {
            float lobeTanHalfAngle = specularLobeTanHalfAngle(ROUGHNESS, 0.3f);
            float worldLobeRadius = hitDist * NoD * lobeTanHalfAngle;
            float lobeRadius = worldLobeRadius / pixelRadiusToWorld(cb.unproject, cb.orthoMode, 1.0f, s.viewZ + hitDist * Dv.w);
            blurRadius = fminf(blurRadius, lobeRadius);
        }

        float2 geomParams = geometryWeightParams(cb.planeDistSensitivity, s.frustumSize, s.Xv, s.Nv);
        float normalParam = normalWeightParam(NLAS, cb.lobeAngleFraction, ROUGHNESS) / fractionScale;
        float2 roughParams = roughnessWeightParams(ROUGHNESS, cb.roughnessFraction * fractionScale);
        float2 hitDistParams = hitDistanceWeightParams(result.w, NLAS);
        float minHitDistWeight = cb.minHitDistanceWeight * fractionScale * smc;
        if (PASS != PRE_PASS) minHitDistWeight *= NLAS;

        float4 scaledRotator = f4(0.0f);
        float3 Tv = f3(0.0f), Bv = f3(0.0f);
        if (SCREEN_SPACE) {
            float2 skew = f2(1.0f);
            if (PASS != PRE_PASS && LOBE == DIFF) {
                skew = lerp(1.0f - fabs2(xy(s.Nv)), f2(1.0f), s.NoV);
                skew /= fmaxf(skew.x, skew.y);
            }
            skew *= rectSizeInv * blurRadius;
            scaledRotator = scaleRotator(s.rotator, skew);
        } else {
            float bentFactor = sqrtf(hdFactor);
            float skewFactor = lerp(0.25f + 0.75f * ROUGHNESS, 1.0f, NoD);
            skewFactor = lerp(skewFactor, 1.0f, NLAS);
            skewFactor = lerp(1.0f, skewFactor, bentFactor);
            float3 bentDv = normalize(lerp(s.Nv, xyz(Dv), bentFactor));
            float worldRadius = pixelRadiusToWorld(cb.unproject, cb.orthoMode, blurRadius, s.viewZ);
            kernelBasis(bentDv, s.Nv, Tv, Bv);
            Tv *= worldRadius * skewFactor;
            Bv *= worldRadius / skewFactor;
        }

        float hitDistForTracking = hitDist == 0.0f ? NRD_INF : hitDist;

#pragma unroll
        for (int n = 0; n < 8; n++) {
            const float ox = kSpecial8X[n], oy = kSpecial8Y[n];

            float2 uv;
            if (SCREEN_SPACE) {
                uv = s.pixelUv + rotate2(scaledRotator, make_float2(ox, oy));
            } else {
                float2 o = rotate2(s.rotator, make_float2(ox, oy));
                float3 p = s.Xv + Tv * o.x + Bv * o.y;
                float4 clip = mulM4(cb.viewToClip, f4(p, 1.0f));
                uv = make_float2(clip.x / clip.w, -(clip.y / clip.w)) * 0.5f + 0.5f;
            }

            float2 muv = mirrorUv(uv);
            // Reference predicate: any( uv != mirrorUv ). For in-screen taps mirrorUv = 1 - ( 1 - uv ) re-rounds uv, so the
            // outcome hangs on the last mantissa bits of uv (see DESIGN.md "chaotic predicates"); the robust variant
            // (debug flag, used by the strict parity tests) asks the intended question: did the tap leave the screen?
            bool mirrored = robustMirrorTest ? (uv.x < 0.0f || uv.y < 0.0f || uv.x >= 1.0f || uv.y >= 1.0f) : (uv.x != muv.x || uv.y != muv.y);
            float w = mirrored ? 1.0f : (n < 4 ? kGaussOuter : kGaussInner);

            float2 posf = muv * rectSize;
            int tx = (int)posf.x, ty = (int)posf.y;

            float zs = unpackViewZ(cb, viewZTex.fetch(tx, ty));  // mirrorUv() < 1 keeps every tap inside the rect: no bounds checks
            float3 Xvs = reconstructViewPosition(make_float2(tx + 0.5f, ty + 0.5f) * rectSizeInv, cb.frustum, zs, cb.orthoMode);

            float materialIDs;
            float4 Ns = unpackNormalRoughness(nrTex.fetchRaw(tx, ty), materialIDs);

            float angle = acosApproxPositive(dot(s.N, xyz(Ns)));
            float NoX = dot(s.Nv, Xvs);

            w *= compareMaterials(s.materialID, materialIDs, MIN_MATERIAL) ? 1.0f : 0.0f;
            w *= nonExponentialWeight(angle, normalParam, 0.0f);
            if (LOBE == SPEC) w *= nonExponentialWeight(Ns.w, roughParams.x, roughParams.y);
            w = applyGeometryWeightLast(cb, w, zs, NoX, geomParams);

            float4 smp = input.fetch(tx, ty);
            smp = w == 0.0f ? f4(0.0f) : smp;

            if (PASS == PRE_PASS && LOBE == SPEC) {
                float hs = smp.w * hitDistanceNormalization(zs, cb.hitDistSettings, Ns.w);
                float geometryWeight = w * s.NoV * (hs != 0.0f ? 1.0f : 0.0f);
                if (rng.next() < geometryWeight) hitDistForTracking = fminf(hitDistForTracking, hs);

                w *= cb.usePrepassNotOnlyForSpecularMotionEstimation;

                float d = length(Xvs - s.Xv) + NRD_EPS;
                float t = hs / (d + hitDist);
                w *= lerp(saturate(t), 1.0f, linearStep(0.5f, 1.0f, ROUGHNESS));
            }

            w *= minHitDistWeight + exponentialWeight(smp.w, hitDistParams.x, hitDistParams.y);

            sum += w;
            result += smp * w;
        }

        result *= positiveRcp(sum);
        if (PASS != PRE_PASS) result.w = hitDist / hitDistScale;
        if (PASS == PRE_PASS && LOBE == SPEC) outSpecHitDistForTracking->store(s.px, s.py, hitDistForTracking == NRD_INF ? 0.0f : hitDistForTracking);
    }

    output.store(s.px, s.py, result);

    if (PASS == POST_BLUR && !temporalStabilization) {
        result.w = cb.returnHistoryLengthInsteadOfOcclusion ? (LOBE == DIFF ? s.data1.x : s.data1.y) : result.w;
        outputCopy->store(s.px, s.py, result);
    }
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// One CTA of 256 threads per 16x16 tile: each thread tests one pixel, the verdict is a block-wide AND.
__global__ void __launch_bounds__(256) reblurClassifyTilesKernel(const __grid_constant__ ReblurConstants cb, const __grid_constant__ ClassifyTilesParams p) {
    int tx = blockIdx.x, ty = blockIdx.y;
    int px = tx * 16 + (threadIdx.x & 15), py = ty * 16 + (threadIdx.x >> 4);
    float viewZ = unpackViewZ(cb, p.inViewZ.load(px, py));
    int allSky = __syncthreads_and(!inDenoisingRange(cb, viewZ));
    if (threadIdx.x == 0) p.outTiles.store(tx, ty, allSky ? 1.0f : 0.0f);
}

__global__ void __launch_bounds__(BLOCK_W* BLOCK_H) reblurPrePassKernel(const __grid_constant__ ReblurConstants cb, const __grid_constant__ PrePassParams p, int flags) {
    const bool robust = (flags & 2) != 0;
    Center s;
    s.px = blockIdx.x * BLOCK_W + threadIdx.x;
    s.py = blockIdx.y * BLOCK_H + threadIdx.y;
    if (p.tiles.load(s.px >> 4, s.py >> 4) != 0.0f || s.px > cb.rectSizeMinusOne[0] || s.py > cb.rectSizeMinusOne[1]) return;
    s.viewZ = unpackViewZ(cb, p.viewZ.load(s.px, s.py));
    if (!inDenoisingRange(cb, s.viewZ)) return;
    setupCenter(cb, s, p.normalRoughness, cb.rotatorPre);
    s.nonLinearAccumSpeed = f2(1.0f / (1.0f + 10.0f));
    s.data1 = f2(0.0f);
    spatialFilter<PRE_PASS, DIFF>(cb, s, p.viewZ, p.normalRoughness, p.inDiff, p.outDiff, nullptr, nullptr, true, robust);
    spatialFilter<PRE_PASS, SPEC>(cb, s, p.viewZ, p.normalRoughness, p.inSpec, p.outSpec, &p.outSpecHitDistForTracking, nullptr, true, robust);
}

// Non-linear accumulation speed with the quad-neighbour smoothing of REBLUR_Blur.cs.hlsl:40-59 (lanes x^1, x^2 of the row)
NRD_DEV float2 quadSmoothedAccumSpeed(const ReblurConstants& cb, float2 data1, float viewZ, bool quads) {
    float2 n = make_float2(advancedNonLinearAccumSpeed(cb, data1.x), advancedNonLinearAccumSpeed(cb, data1.y));
    if (!inDenoisingRange(cb, viewZ)) n = f2(0.0f);
    if (quads) {
        float2 d10 = make_float2(__shfl_xor_sync(0xFFFFFFFFu, n.x, 1), __shfl_xor_sync(0xFFFFFFFFu, n.y, 1));
        float2 d01 = make_float2(__shfl_xor_sync(0xFFFFFFFFu, n.x, 2), __shfl_xor_sync(0xFFFFFFFFu, n.y, 2));
        float2 avg = (d10 + d01 + n) / 3.0f;
        n = min2(n, avg);
    }
    return n;
}

__global__ void __launch_bounds__(BLOCK_W* BLOCK_H) reblurBlurKernel(const __grid_constant__ ReblurConstants cb, const __grid_constant__ BlurParams p, int flags) {
    const bool quads = (flags & 1) != 0, robust = (flags & 2) != 0;
    Center s;
    s.px = blockIdx.x * BLOCK_W + threadIdx.x;
    s.py = blockIdx.y * BLOCK_H + threadIdx.y;

    // viewZ (sky included) is copied for the next pass and the next frame
    float viewZpacked = p.viewZ.load(s.px, s.py);
    p.outViewZ.store(s.px, s.py, viewZpacked);

    // No lane leaves before the quad exchange; lanes of sky tiles / outside the rect only feed their own quads
    bool skyTile = p.tiles.load(s.px >> 4, s.py >> 4) != 0.0f;
    s.data1 = unpackData1(p.data1.load(s.px, s.py));
    s.viewZ = unpackViewZ(cb, viewZpacked);
    s.nonLinearAccumSpeed = quadSmoothedAccumSpeed(cb, s.data1, s.viewZ, quads);
    if (skyTile || !inDenoisingRange(cb, s.viewZ) || s.px > cb.rectSizeMinusOne[0] || s.py > cb.rectSizeMinusOne[1]) return;

    setupCenter(cb, s, p.normalRoughness, cb.rotator);
    spatialFilter<BLUR, DIFF>(cb, s, p.viewZ, p.normalRoughness, p.inDiff, p.outDiff, nullptr, nullptr, true, robust);
    spatialFilter<BLUR, SPEC>(cb, s, p.viewZ, p.normalRoughness, p.inSpec, p.outSpec, nullptr, nullptr, true, robust);
}

template <bool TEMPORAL_STABILIZATION>
__global__ void __launch_bounds__(BLOCK_W* BLOCK_H) reblurPostBlurKernel(const __grid_constant__ ReblurConstants cb, const __grid_constant__ PostBlurParams p, int flags) {
    const bool quads = (flags & 1) != 0, robust = (flags & 2) != 0;
    Center s;
    s.px = blockIdx.x * BLOCK_W + threadIdx.x;
    s.py = blockIdx.y * BLOCK_H + threadIdx.y;

    bool skyTile = p.tiles.load(s.px >> 4, s.py >> 4) != 0.0f;
    s.data1 = unpackData1(p.data1.load(s.px, s.py));
    s.viewZ = unpackViewZ(cb, p.viewZ.load(s.px, s.py));
    s.nonLinearAccumSpeed = quadSmoothedAccumSpeed(cb, s.data1, s.viewZ, quads);
    if (skyTile || !inDenoisingRange(cb, s.viewZ) || s.px > cb.rectSizeMinusOne[0] || s.py > cb.rectSizeMinusOne[1]) return;

    setupCenter(cb, s, p.normalRoughness, cb.rotatorPost);

    p.outNormalRoughness.storeRaw(s.px, s.py, p.normalRoughness.loadRaw(s.px, s.py));
    if (!TEMPORAL_STABILIZATION) p.outInternalData.store(s.px, s.py, packInternalData(cb, s.data1.x, s.data1.y, s.materialID));

    spatialFilter<POST_BLUR, DIFF>(cb, s, p.viewZ, p.normalRoughness, p.inDiff, p.outDiff, nullptr, &p.outDiffCopy, TEMPORAL_STABILIZATION, robust);
    spatialFilter<POST_BLUR, SPEC>(cb, s, p.viewZ, p.normalRoughness, p.inSpec, p.outSpec, nullptr, &p.outSpecCopy, TEMPORAL_STABILIZATION, robust);
}

// ---------------------------------------------------------------------------------------------------------------
// Host launchers (called by the executor)
// ---------------------------------------------------------------------------------------------------------------
static dim3 pixelGrid(const ReblurConstants& cb) { return dim3((cb.rectSizeMinusOne[0] + BLOCK_W) / BLOCK_W, (cb.rectSizeMinusOne[1] + BLOCK_H) / BLOCK_H); }

void launchReblurClassifyTiles(const ReblurConstants& cb, const ClassifyTilesParams& p, cudaStream_t stream) {
    dim3 grid((cb.rectSizeMinusOne[0] + 16) / 16, (cb.rectSizeMinusOne[1] + 16) / 16);
    reblurClassifyTilesKernel<<<grid, 256, 0, stream>>>(cb, p);
}
void launchReblurPrePass(const ReblurConstants& cb, const PrePassParams& p, int flags, cudaStream_t stream) {
    reblurPrePassKernel<<<pixelGrid(cb), dim3(BLOCK_W, BLOCK_H), 0, stream>>>(cb, p, flags);
}
void launchReblurBlur(const ReblurConstants& cb, const BlurParams& p, int flags, cudaStream_t stream) {
    reblurBlurKernel<<<pixelGrid(cb), dim3(BLOCK_W, BLOCK_H), 0, stream>>>(cb, p, flags);
}
void launchReblurPostBlur(const ReblurConstants& cb, const PostBlurParams& p, bool temporalStabilization, int flags, cudaStream_t stream) {
    if (temporalStabilization)
        reblurPostBlurKernel<true><<<pixelGrid(cb), dim3(BLOCK_W, BLOCK_H), 0, stream>>>(cb, p, flags);
    else
        reblurPostBlurKernel<false><<<pixelGrid(cb), dim3(BLOCK_W, BLOCK_H), 0, stream>>>(cb, p, flags);
}

}  // namespace nrdk
