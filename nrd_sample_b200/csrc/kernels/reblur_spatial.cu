// REBLUR spatial passes on sm_100a: ClassifyTiles, PrePass, Blur, PostBlur (NRD_MODE=RADIANCE; NRD_SIGNAL = DIFF / SPEC / BOTH as a template
// parameter: REBLUR_DIFFUSE and REBLUR_SPECULAR run the same code with the other lobe compiled out).
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
//
// The kernels are issue-bound on fp32 arithmetic (ncu: ~80 % issue-slot utilisation, DRAM < 5 %), so the tap loop is
// written for Blackwell's packed fp32x2 pipe: taps n and n+4 (outer / inner ring of g_Special8) advance in lock step as
// the two lanes of FFMA2 / FMUL2 / FADD2 (pairmath.cuh); only loads, conversions, MUFU and min/max/abs stay scalar.
#include "pairmath.cuh"
#include "reblur_common.cuh"

namespace nrdk {

namespace {

enum { PRE_PASS = 0, BLUR = 1, POST_BLUR = 2 };
enum { DIFF = 0, SPEC = 1 };

constexpr int BLOCK_W = 32, BLOCK_H = 8;
#ifndef SPATIAL_MIN_BLOCKS
#    define SPATIAL_MIN_BLOCKS 4  // resident CTAs per SM the register allocator must leave room for: 64 regs + a 32-48 B spill beats
                                 // 80 regs (3 CTAs) and 124 regs (2 CTAs) by 5 % / 13 % on B200 (profiles/README.md)
#endif

// g_Special8 (Common.hlsli:207-218) as compile-time immediates: the pair loop is fully unrolled, offsets fold into the
// instruction stream and the per-tap gaussian weight exp(-0.66 z^2), z in {1, 0.5}, becomes a literal.
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

// Checkerboard resolve state of the pre-pass (REBLUR_PrePass.cs.hlsl:52-67): parity of the pixel, half-width x of its row
// neighbours and their disocclusion weights
struct Resolve {
    uint32_t checkerboard;
    int x0, x1;
    float2 wc;
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

// Math::SmoothStep( 1, 0, |y| ) for a pair
NRD_DEV P2 nonExponentialWeight2(P2 y) {
    P2 s = oneMinusAbsSat2(y);
    return (s * s) * fma2(s, -2.0f, 3.0f);
}
// ExpApprox( -3 |y| ) = 1 / ( x^2 - x + 1 ), x = -3 |y|
NRD_DEV P2 exponentialWeight2(P2 y) {
    P2 x = absMul2(y, -3.0f);
    return rcp2(fma2(x, x, fma2(x, -1.0f, 1.0f)));
}

// One lobe of one spatial pass for one pixel
// CB: checkerboard input (pre-pass only): `input` is half width, holds the pixels whose parity equals `cbMode` this frame
// SH ( NRD_MODE = SH ): a second RGBA16F per lobe rides along with the weights of the first ( REBLUR_Common_SpatialFilter.hlsli:67-69, 268-280, 307-335 )
struct ShIo { const TexRGBA16F *in, *out, *outCopy; };
// PROBE ( NRDCU_FLAG_PROBE_MIRROR, tests only ): counts the taps and how many of them took the "mirrored" branch of the Gaussian weight
// ( REBLUR_Common_SpatialFilter.hlsli:198 ) into g_mirrorProbe — the predicate is decided by the last mantissa bit of the tap position, so
// parity with the reference is stated on its RATE ( tests/test_parity_at_baseline_sizes_gpu.py ). Compiled out of every other instantiation.
__device__ unsigned long long g_mirrorProbe[6][2];   // [ pass * 2 + lobe ][ taps, mirrored ]
// MODE: NRD_MODE of the permutation ( reblur_common.cuh ). OCCLUSION / DO read and write their signal through the format-polymorphic Sig< MODE >, keep
// minHitDistWeight unscaled and filter the hit distance like any other channel ( REBLUR_Common_SpatialFilter.hlsli:140-142, 283-285, 327-329 ).
template <int PASS, int LOBE, bool CB = false, int MODE = MODE_RADIANCE, bool PROBE = false>
NRD_DEV void spatialFilter(const ReblurConstants& cb, const Center& s, const TexGeom& geom, const TexNR& nrTex, const TexRGBA16F& input,
                           const TexRGBA16F& output, const TexR16F* outSpecHitDistForTracking, const TexRGBA16F* outputCopy, bool temporalStabilization, bool robustMirrorTest,
                           const Resolve* resolve = nullptr, ShIo shIo = ShIo()) {
    static_assert(!CB || PASS == PRE_PASS, "only the pre-pass reads checkerboarded input");
    constexpr bool SH = MODE == MODE_SH, FIXED = Sig<MODE>::FIXED;
    const uint32_t cbMode = CB ? (LOBE == DIFF ? cb.diffCheckerboard : cb.specCheckerboard) : 2u;
    const float ROUGHNESS = LOBE == DIFF ? 1.0f : s.roughness;
    const float NLAS = LOBE == DIFF ? s.nonLinearAccumSpeed.x : s.nonLinearAccumSpeed.y;
    const float MIN_MATERIAL = LOBE == DIFF ? cb.diffMinMaterial : cb.specMinMaterial;
    const float MAX_BLUR_RADIUS = PASS == PRE_PASS ? (LOBE == DIFF ? cb.diffPrepassBlurRadius : cb.specPrepassBlurRadius) : cb.maxBlurRadius;
    constexpr bool SCREEN_SPACE = PASS == PRE_PASS || LOBE == DIFF;
    const float2 rectSize = make_float2(cb.rectSize[0], cb.rectSize[1]);
    const float2 rectSizeInv = make_float2(cb.rectSizeInv[0], cb.rectSizeInv[1]);

    float sum = 1.0f;
    float4 result = Sig<MODE>::load(input, CB ? s.px >> 1 : s.px, s.py);
    float4 resultSh = f4(0.0f);
    if constexpr (SH) resultSh = shIo.in->load(CB ? s.px >> 1 : s.px, s.py);
    if (CB && resolve->checkerboard != cbMode) {
        sum = 0.0f;
        result = f4(0.0f);
        resultSh = f4(0.0f);
    }

    if (PASS != PRE_PASS || MAX_BLUR_RADIUS != 0.0f) {
        // Stochastic tracking decisions of the specular pre-pass consume one random number per tap, in tap order
        float rnd[8];
        if (PASS == PRE_PASS && LOBE == SPEC) {
            Rng rng;
            rng.init((uint32_t)s.px, (uint32_t)s.py, cb.frameIndex);
#pragma unroll
            for (int n = 0; n < 8; n++) rnd[n] = rng.next();
        }

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

        const float2 geomParams = geometryWeightParams(cb.planeDistSensitivity, s.frustumSize, s.Xv, s.Nv);
        const float normalParam = normalWeightParam(NLAS, cb.lobeAngleFraction, ROUGHNESS) / fractionScale;
        const float2 roughParams = roughnessWeightParams(ROUGHNESS, cb.roughnessFraction * fractionScale);
        const float2 hitDistParams = hitDistanceWeightParams(result.w, NLAS);
        float minHitDistWeight = cb.minHitDistanceWeight * fractionScale * smc;
        if (PASS != PRE_PASS && FIXED) minHitDistWeight *= NLAS;
        // CompareMaterials( m0, m, minm ) = max( m0, minm ) == max( m, minm ) on the 2-bit IDs k0, k in 0..3 is
        // ( k == k0 ) || max( k, k0 ) <= minm: integer compares, and no work at all when minm >= 3 (the default is 4)
        const uint32_t centerK = (uint32_t)(s.materialID + 0.5f);
        const bool materialsAlwaysMatch = MIN_MATERIAL >= 3.0f;

        // Tap placement. Screen space: uv = pixelUv + R * offset with the blur-radius-scaled rotator R. World space:
        // clip = M * ( Xv + T o.x + B o.y ) is expanded once per pixel into C + A o.x + B' o.y (rows x, y, w of M only).
        float4 scaledRotator = f4(0.0f);
        float3 clipC = f3(0.0f), clipA = f3(0.0f), clipB = f3(0.0f);
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
            float3 Tv, Bv;
            kernelBasis(bentDv, s.Nv, Tv, Bv);
            Tv *= worldRadius * skewFactor;
            Bv *= worldRadius / skewFactor;
            const Mat4& M = cb.viewToClip;
            auto rows = [&](float3 v, float w) {
                return make_float3(M.m[0] * v.x + M.m[4] * v.y + M.m[8] * v.z + M.m[12] * w, M.m[1] * v.x + M.m[5] * v.y + M.m[9] * v.z + M.m[13] * w,
                                   M.m[3] * v.x + M.m[7] * v.y + M.m[11] * v.z + M.m[15] * w);
            };
            clipC = rows(s.Xv, 1.0f);
            clipA = rows(Tv, 0.0f);
            clipB = rows(Bv, 0.0f);
        }

        // view-space ray through the centre of texel (tx, ty): ray = ( tx + 0.5 ) / rect * frustum.zw + frustum.xy, folded into one FMA
        const float rayMulX = rectSizeInv.x * cb.frustum[2], rayAddX = 0.5f * rectSizeInv.x * cb.frustum[2] + cb.frustum[0];
        const float rayMulY = rectSizeInv.y * cb.frustum[3], rayAddY = 0.5f * rectSizeInv.y * cb.frustum[3] + cb.frustum[1];
        const bool perspective = cb.orthoMode == 0.0f;

        float hitDistForTracking = hitDist == 0.0f ? NRD_INF : hitDist;
        const float mirrorEps = robustMirrorTest ? 1e-6f : 0.0f;
        unsigned probeMirrored = 0u;
        P2 sum2(0.0f);
        P2 accX(0.0f), accY(0.0f), accZ(0.0f), accW(0.0f);
        P2 shX(0.0f), shY(0.0f), shZ(0.0f), shW(0.0f);

#pragma unroll
        for (int pair = 0; pair < 4; pair++) {
            const int na = pair, nb = pair + 4;
            const P2 OX(kSpecial8X[na], kSpecial8X[nb]), OY(kSpecial8Y[na], kSpecial8Y[nb]);

            P2 ux, uy;
            if (SCREEN_SPACE) {
                ux = fma2(OY, scaledRotator.y, fma2(OX, scaledRotator.x, P2(s.pixelUv.x)));
                uy = fma2(OY, scaledRotator.w, fma2(OX, scaledRotator.z, P2(s.pixelUv.y)));
            } else {
                // rotated unit offsets are uniform per tap
                P2 ox = fma2(OY, s.rotator.y, OX * s.rotator.x), oy = fma2(OY, s.rotator.w, OX * s.rotator.z);
                P2 cx = fma2(oy, clipB.x, fma2(ox, clipA.x, P2(clipC.x)));
                P2 cy = fma2(oy, clipB.y, fma2(ox, clipA.y, P2(clipC.y)));
                P2 cw = fma2(oy, clipB.z, fma2(ox, clipA.z, P2(clipC.z)));
                // Geometry::GetScreenUv ( ml.hlsli:659-672 ): clip.xy / clip.w * ( 0.5, -0.5 ) + 0.5. The product is ROUNDED before 0.5 is added, as in the
                // reference: a tap at uv < 0.25 is then an exact sum with trailing zero mantissa bits, which is what the reference's mirror predicate
                // ( uv != MirrorUv( uv ), decided by those bits ) sees. A fused multiply-add here took the "mirrored" branch 12 % more often than the
                // reference's shaders at 1080p ( 0.42 against 0.31 in the specular lobes of blur / post-blur; tests/test_parity_at_baseline_sizes_gpu.py
                // measures the rate of both engines per pass and lobe ).
                P2 iw = rcp2(cw) * 0.5f;
                ux = mulThenAdd2(cx, iw, 0.5f);
                uy = mulThenAdd2(cy * -1.0f, iw, 0.5f);
            }

            // MirrorUv (Common.hlsli:312-318): 1 - | 1 - frac( uv / 2 ) * 2 |, capped below 1
            P2 hx = ux * 0.5f, hy = uy * 0.5f;
            P2 mx = min2(oneMinusAbsSat2(fma2(hx - floor2(hx), -2.0f, 1.0f)), 0.99999f);
            P2 my = min2(oneMinusAbsSat2(fma2(hy - floor2(hy), -2.0f, 1.0f)), 0.99999f);
            // Reference predicate: any( uv != mirrorUv ). For in-screen taps mirrorUv = 1 - ( 1 - uv ) re-rounds uv, so the outcome hangs on the last
            // mantissa bits of uv (DESIGN.md "chaotic predicates"); the robust variant (debug flag, used by the strict parity tests) asks the intended
            // question — did mirroring MOVE the tap? — as | uv - mirrorUv | > 1e-6 ( re-rounding moves it by <= 2^-24 ). One code path for both:
            // the threshold is 0 for the reference's predicate.
            const P2 dux = ux - mx, duy = uy - my;
            const bool mirA = fabsf(dux.a()) > mirrorEps || fabsf(duy.a()) > mirrorEps;
            const bool mirB = fabsf(dux.b()) > mirrorEps || fabsf(duy.b()) > mirrorEps;
            P2 w(mirA ? 1.0f : kGaussOuter, mirB ? 1.0f : kGaussInner);
            if constexpr (PROBE) probeMirrored += (mirA ? 1u : 0u) + (mirB ? 1u : 0u);

            // texel coordinates: mirrorUv() < 1 keeps every tap inside the rect, so fetches need no bounds checks. F2I.FLOOR gives the texel, the
            // float copy the view ray needs comes back through I2FP ( ALU pipe ) instead of a second rounding on the quarter-rate XU pipe
            const P2 sx = mx * rectSize.x, sy = my * rectSize.y;
            int txa = __float2int_rd(sx.a()), txb = __float2int_rd(sx.b());
            const int tya = __float2int_rd(sy.a()), tyb = __float2int_rd(sy.b());
            P2 fx((float)txa, (float)txb);
            const P2 fy((float)tya, (float)tyb);
            int ixa = txa, ixb = txb;  // x in the (possibly half-width) input
            if (CB) {
                // Move to a pixel that was traced this frame: taps n = pair (even / odd) and n + 4 share the shift direction
                const int shift = (pair & 1) == 0 ? -1 : 1;
                txa += (((uint32_t)(txa ^ tya) ^ cb.frameIndex) & 1u) != cbMode ? shift : 0;
                txb += (((uint32_t)(txb ^ tyb) ^ cb.frameIndex) & 1u) != cbMode ? shift : 0;
                ixa = txa >> 1;
                ixb = txb >> 1;
                // a tap pushed off the rect is invalid; it fetches the clamped texel and gets zero weight
                if (txa < 0 || txa > cb.rectSizeMinusOne[0]) w = P2(0.0f, w.b());
                if (txb < 0 || txb > cb.rectSizeMinusOne[0]) w = P2(w.a(), 0.0f);
                const int cxa = clampi(txa, 0, cb.rectSizeMinusOne[0]), cxb = clampi(txb, 0, cb.rectSizeMinusOne[0]);
                fx = P2((float)txa, (float)txb);
                txa = cxa;
                txb = cxb;
                ixa = clampi(ixa, 0, input.w - 1);
                ixb = clampi(ixb, 0, input.w - 1);
            }

            // One 128-bit fetch per tap from the geometry plane ( reblurGeometryPlaneKernel ): { world-space normal, |viewZ| } decoded ONCE per texel
            // per frame instead of once per tap — 48 taps per pixel and frame read it back ( NRD.hlsli:387-400, 656-684; Common.hlsli:261 )
            const float4 gA = geom.fetch(txa, tya), gB = geom.fetch(txb, tyb);
            uint2 rawA = make_uint2(0u, 0u), rawB = make_uint2(0u, 0u);
            float4 smpA = f4(0.0f), smpB = f4(0.0f);
            if constexpr (FIXED) {
                rawA = input.fetchRaw(ixa, tya);
                rawB = input.fetchRaw(ixb, tyb);
            } else {
                smpA = Sig<MODE>::fetch(input, ixa, tya);
                smpB = Sig<MODE>::fetch(input, ixb, tyb);
            }
            uint2 rawShA = make_uint2(0u, 0u), rawShB = make_uint2(0u, 0u);
            if constexpr (SH) {
                rawShA = shIo.in->fetchRaw(ixa, tya);
                rawShB = shIo.in->fetchRaw(ixb, tyb);
            }

            const P2 zs(gA.w, gB.w);
            P2 rx = fma2(fx, rayMulX, rayAddX), ry = fma2(fy, rayMulY, rayAddY);
            P2 sxy = perspective ? zs : P2(cb.orthoMode);
            // dot( N, Ns ): scalar FFMAs straight from the two fetched quads into one register pair ( pairing x / y / z across the taps first would cost moves )
            const P2 cosa(fmaf(gA.z, s.N.z, fmaf(gA.y, s.N.y, gA.x * s.N.x)), fmaf(gB.z, s.N.z, fmaf(gB.y, s.N.y, gB.x * s.N.x)));

            // roughness ( specular lobe ) and material ID still come from the raw 10:10:10:2 texel: 3 instructions per tap, only where they are used
            P2 roughS(0.0f);
            bool matchA = true, matchB = true;
            if (LOBE == SPEC || !materialsAlwaysMatch) {
                const uint32_t nrA = nrTex.fetchRaw(txa, tya), nrB = nrTex.fetchRaw(txb, tyb);
                if (LOBE == SPEC) {
                    const P2 pz10((float)((nrA >> 20) & 1023u), (float)((nrB >> 20) & 1023u));
                    roughS = abs2(fma2(pz10, 2.0f / 1023.0f, -1.0f));
                }
                if (!materialsAlwaysMatch) {
                    const uint32_t ka = nrA >> 30, kb = nrB >> 30;
                    matchA = ka == centerK || (float)max(ka, centerK) <= MIN_MATERIAL;
                    matchB = kb == centerK || (float)max(kb, centerK) <= MIN_MATERIAL;
                }
            }

            // Math::AcosApproxPositive
            P2 cs = sat2(cosa);
            P2 angle = fma2(cs, 1.399331f - 1.567589f, 1.567589f) * sqrt2(satOneMinus2(cosa));

            if (!materialsAlwaysMatch) w = sel2(matchA, matchB, w, P2(0.0f));
            w = w * nonExponentialWeight2(angle * normalParam);
            if (LOBE == SPEC) w = w * nonExponentialWeight2(fma2(roughS, roughParams.x, roughParams.y));
            // plane distance: NoX = dot( Nv, Xvs ), Xvs = ( ray.xy * sxy, zs )
            P2 NoX = fma2(zs, s.Nv.z, sxy * fma2(ry, s.Nv.y, rx * s.Nv.x));
            w = w * nonExponentialWeight2(fma2(NoX, geomParams.x, geomParams.y));
            // ( taps beyond the denoising range: zs = +INF in the plane, the plane-distance weight above is 0 for them )

            // Denanify, on the packed fp16 words (2 selects per tap instead of 4)
            if constexpr (FIXED) {
                if (w.a() == 0.0f) rawA = make_uint2(0u, 0u);
                if (w.b() == 0.0f) rawB = make_uint2(0u, 0u);
                smpA = TexRGBA16F::decode(rawA);
                smpB = TexRGBA16F::decode(rawB);
            } else {
                if (w.a() == 0.0f) smpA = f4(0.0f);
                if (w.b() == 0.0f) smpB = f4(0.0f);
            }
            P2 sw(smpA.w, smpB.w);

            if (PASS == PRE_PASS && LOBE == SPEC) {
                // hit distance for tracking: stochastic min over taps weighted by their geometry weight
                P2 smcS(specMagicCurve(roughS.a(), 0.5f), specMagicCurve(roughS.b(), 0.5f));
                P2 hs = sw * (fma2(zs, cb.hitDistSettings[1], cb.hitDistSettings[0]) * fma2(smcS, 1.0f - cb.hitDistSettings[2], cb.hitDistSettings[2]));
                P2 geometryWeight = w * s.NoV;
                if (hs.a() != 0.0f && rnd[na] < geometryWeight.a()) hitDistForTracking = fminf(hitDistForTracking, hs.a());
                if (hs.b() != 0.0f && rnd[nb] < geometryWeight.b()) hitDistForTracking = fminf(hitDistForTracking, hs.b());

                w = w * cb.usePrepassNotOnlyForSpecularMotionEstimation;

                P2 dx = fma2(rx, sxy, -s.Xv.x), dy = fma2(ry, sxy, -s.Xv.y), dz = zs - s.Xv.z;
                P2 d = sqrt2(fma2(dz, dz, fma2(dy, dy, dx * dx))) + NRD_EPS;
                P2 tt = mulSat2(hs, rcp2(d + hitDist));
                float k = linearStep(0.5f, 1.0f, ROUGHNESS);
                w = w * fma2(tt, 1.0f - k, k);  // lerp( saturate( t ), 1, k ) = t ( 1 - k ) + k
            }

            w = w * (exponentialWeight2(fma2(sw, hitDistParams.x, hitDistParams.y)) + minHitDistWeight);

            sum2 = sum2 + w;
            accX = fma2(P2(smpA.x, smpB.x), w, accX);
            accY = fma2(P2(smpA.y, smpB.y), w, accY);
            accZ = fma2(P2(smpA.z, smpB.z), w, accZ);
            accW = fma2(sw, w, accW);
            if constexpr (SH) {
                if (w.a() == 0.0f) rawShA = make_uint2(0u, 0u);  // Denanify by the FINAL weight ( :268-272 )
                if (w.b() == 0.0f) rawShB = make_uint2(0u, 0u);
                const float4 a4 = TexRGBA16F::decode(rawShA), b4 = TexRGBA16F::decode(rawShB);
                shX = fma2(P2(a4.x, b4.x), w, shX);
                shY = fma2(P2(a4.y, b4.y), w, shY);
                shZ = fma2(P2(a4.z, b4.z), w, shZ);
                shW = fma2(P2(a4.w, b4.w), w, shW);
            }
        }

        if constexpr (PROBE) {
            const unsigned active = __activemask();
            const unsigned mirrored = __reduce_add_sync(active, probeMirrored);
            if ((threadIdx.x & 31) == __ffs(active) - 1) {
                atomicAdd(&g_mirrorProbe[PASS * 2 + LOBE][0], 8ull * __popc(active));
                atomicAdd(&g_mirrorProbe[PASS * 2 + LOBE][1], (unsigned long long)mirrored);
            }
        }
        sum += sum2.a() + sum2.b();
        result += make_float4(accX.a() + accX.b(), accY.a() + accY.b(), accZ.a() + accZ.b(), accW.a() + accW.b());
        result *= positiveRcp(sum);
        if constexpr (SH) {
            resultSh += make_float4(shX.a() + shX.b(), shY.a() + shY.b(), shZ.a() + shZ.b(), shW.a() + shW.b());
            resultSh *= positiveRcp(sum);
        }
        if (PASS != PRE_PASS && FIXED) result.w = hitDist / hitDistScale;
        if (PASS == PRE_PASS && LOBE == SPEC) outSpecHitDistForTracking->store(s.px, s.py, hitDistForTracking == NRD_INF ? 0.0f : hitDistForTracking);
    }

    // Checkerboard resolve ( if the pre-pass found nothing ): the traced row neighbours, REBLUR_Common_SpatialFilter.hlsli:294-316
    if (CB && sum == 0.0f) {
        float4 s0 = Sig<MODE>::load(input, resolve->x0, s.py), s1 = Sig<MODE>::load(input, resolve->x1, s.py);
        if (resolve->wc.x == 0.0f) s0 = f4(0.0f);
        if (resolve->wc.y == 0.0f) s1 = f4(0.0f);
        result = s0 * resolve->wc.x + s1 * resolve->wc.y;
        if constexpr (SH) {
            float4 sh0 = shIo.in->load(resolve->x0, s.py), sh1 = shIo.in->load(resolve->x1, s.py);
            if (resolve->wc.x == 0.0f) sh0 = f4(0.0f);
            if (resolve->wc.y == 0.0f) sh1 = f4(0.0f);
            resultSh = sh0 * resolve->wc.x + sh1 * resolve->wc.y;
        }
    }

    Sig<MODE>::store(output, s.px, s.py, result);
    if constexpr (SH) shIo.out->store(s.px, s.py, resultSh);

    if (PASS == POST_BLUR && !temporalStabilization) {
        if (FIXED) result.w = cb.returnHistoryLengthInsteadOfOcclusion ? (LOBE == DIFF ? s.data1.x : s.data1.y) : result.w;
        Sig<MODE>::store(*outputCopy, s.px, s.py, result);
        if constexpr (SH) shIo.outCopy->store(s.px, s.py, resultSh);
    }
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// One warp per 16x16 tile, 8 tiles per CTA ( tileIsSkyWarp, common.cuh )
__global__ void __launch_bounds__(256) reblurClassifyTilesKernel(const __grid_constant__ ReblurConstants cb, const __grid_constant__ ClassifyTilesParams p, int ctaY0, int tilesW) {
    pdlEntry();
    const int tx = blockIdx.x * 8 + (threadIdx.x >> 5), ty = ctaY0 + blockIdx.y;
    if (tx >= tilesW) return;
    const bool allSky = tileIsSkyWarp(p.inViewZ, tx, ty, [&](float z) { return !inDenoisingRange(cb, unpackViewZ(cb, z)); });
    if ((threadIdx.x & 31) == 0) p.outTiles.store(tx, ty, allSky ? 1.0f : 0.0f);
}

// Decodes IN_NORMAL_ROUGHNESS + IN_VIEWZ into the geometry plane, one thread per texel. Same arithmetic as the centre set-up of the passes
// ( unpackNormalRoughness, |viewZ * scale| ), so a tap that lands on the centre texel sees the centre's own normal.
__global__ void __launch_bounds__(256) reblurGeometryPlaneKernel(const __grid_constant__ ReblurConstants cb, const __grid_constant__ GeometryPlaneParams p, int row0, int row1) {
    pdlEntry();
    const int px = blockIdx.x * 32 + (threadIdx.x & 31), py = row0 + blockIdx.y * 8 + (threadIdx.x >> 5);
    if (px > cb.rectSizeMinusOne[0] || py >= row1) return;
    const float4 nr = unpackNormalRoughness(p.normalRoughness.fetchRaw(px, py));
    // texels outside the denoising range carry +INF: every tap weight includes the plane-distance term sat( 1 - | NoX a + b | ) with NoX proportional
    // to this value, so such a tap weighs exactly 0 ( INF or NaN -> FADD.SAT -> 0 ) without a per-tap range compare ( Common.hlsli:567-572 )
    const float zs = fabsf(p.viewZ.fetch(px, py)) * fabsf(cb.viewZScale);
    p.out.store(px, py, make_float4(nr.x, nr.y, nr.z, zs < cb.denoisingRange ? zs : __int_as_float(0x7F800000)));
}

template <bool CB, int SIGNAL, int MODE, bool PROBE = false>
__global__ void __launch_bounds__(BLOCK_W* BLOCK_H, SPATIAL_MIN_BLOCKS) reblurPrePassKernel(const __grid_constant__ ReblurConstants cb, const __grid_constant__ PrePassParams p, int flags, int ctaY0) {
    pdlEntry();
    const bool robust = (flags & 2) != 0;
    Center s;
    const int2 cta = ctaTile<0>(ctaY0);
    s.px = cta.x * BLOCK_W + threadIdx.x;
    s.py = cta.y * BLOCK_H + threadIdx.y;
    if (p.tiles.load(s.px >> 4, s.py >> 4) != 0.0f || s.px > cb.rectSizeMinusOne[0] || s.py > cb.rectSizeMinusOne[1]) return;
    s.viewZ = unpackViewZ(cb, p.viewZ.load(s.px, s.py));
    if (!inDenoisingRange(cb, s.viewZ)) return;
    setupCenter(cb, s, p.normalRoughness, cb.rotatorPre);
    s.nonLinearAccumSpeed = f2(1.0f / (1.0f + 10.0f));
    s.data1 = f2(0.0f);
    Resolve r = {};
    if (CB) {
        r.checkerboard = ((uint32_t)(s.px ^ s.py) ^ cb.frameIndex) & 1u;
        const int x0 = max(s.px - 1, 0), x1 = min(s.px + 1, cb.rectSizeMinusOne[0]);
        const float viewZ0 = unpackViewZ(cb, p.viewZ.load(x0, s.py)), viewZ1 = unpackViewZ(cb, p.viewZ.load(x1, s.py));
        const float threshold = disocclusionThresholdAt(0.02f, s.frustumSize, s.NoV);  // NRD_DISOCCLUSION_THRESHOLD
        float2 wc = make_float2(threshold >= fabsf(viewZ0 - s.viewZ) ? 1.0f : 0.0f, threshold >= fabsf(viewZ1 - s.viewZ) ? 1.0f : 0.0f);
        if (!inDenoisingRange(cb, viewZ0) || s.px < 1) wc.x = 0.0f;
        if (!inDenoisingRange(cb, viewZ1) || s.px >= cb.rectSizeMinusOne[0]) wc.y = 0.0f;
        r.wc = wc * positiveRcp(wc.x + wc.y);
        r.x0 = x0 >> 1;
        r.x1 = x1 >> 1;
    }
    if constexpr ((SIGNAL & SIGNAL_DIFF) != 0) spatialFilter<PRE_PASS, DIFF, CB, MODE, PROBE>(cb, s, p.geom, p.normalRoughness, p.inDiff, p.outDiff, nullptr, nullptr, true, robust, &r, ShIo{&p.inDiffSh, &p.outDiffSh, nullptr});
    if constexpr ((SIGNAL & SIGNAL_SPEC) != 0) spatialFilter<PRE_PASS, SPEC, CB, MODE, PROBE>(cb, s, p.geom, p.normalRoughness, p.inSpec, p.outSpec, &p.outSpecHitDistForTracking, nullptr, true, robust, &r, ShIo{&p.inSpecSh, &p.outSpecSh, nullptr});
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

template <int SIGNAL, int MODE, bool PROBE = false>
__global__ void __launch_bounds__(BLOCK_W* BLOCK_H, SPATIAL_MIN_BLOCKS) reblurBlurKernel(const __grid_constant__ ReblurConstants cb, const __grid_constant__ BlurParams p, int flags, int ctaY0) {
    pdlEntry();
    const bool quads = (flags & 1) != 0, robust = (flags & 2) != 0;
    Center s;
    const int2 cta = ctaTile<3>(ctaY0);
    s.px = cta.x * BLOCK_W + threadIdx.x;
    s.py = cta.y * BLOCK_H + threadIdx.y;

    // viewZ (sky included) is copied for the next pass and the next frame
    // ( by every thread of the reference's 8x16 groups: up to the rect rounded up to 8 x 16; with dynamic resolution the texels beyond belong to nobody )
    float viewZpacked = p.viewZ.load(s.px, s.py);
    if (s.px < ((cb.rectSizeMinusOne[0] + 8) & ~7) && s.py < ((cb.rectSizeMinusOne[1] + 16) & ~15)) p.outViewZ.store(s.px, s.py, viewZpacked);

    // No lane leaves before the quad exchange; lanes of sky tiles / outside the rect only feed their own quads
    bool skyTile = p.tiles.load(s.px >> 4, s.py >> 4) != 0.0f;
    s.data1 = loadData1<SIGNAL>(p, s.px, s.py);
    s.viewZ = unpackViewZ(cb, viewZpacked);
    s.nonLinearAccumSpeed = quadSmoothedAccumSpeed(cb, s.data1, s.viewZ, quads);
    if (skyTile || !inDenoisingRange(cb, s.viewZ) || s.px > cb.rectSizeMinusOne[0] || s.py > cb.rectSizeMinusOne[1]) return;

    setupCenter(cb, s, p.normalRoughness, cb.rotator);
    if constexpr ((SIGNAL & SIGNAL_DIFF) != 0) spatialFilter<BLUR, DIFF, false, MODE, PROBE>(cb, s, p.geom, p.normalRoughness, p.inDiff, p.outDiff, nullptr, nullptr, true, robust, nullptr, ShIo{&p.inDiffSh, &p.outDiffSh, nullptr});
    if constexpr ((SIGNAL & SIGNAL_SPEC) != 0) spatialFilter<BLUR, SPEC, false, MODE, PROBE>(cb, s, p.geom, p.normalRoughness, p.inSpec, p.outSpec, nullptr, nullptr, true, robust, nullptr, ShIo{&p.inSpecSh, &p.outSpecSh, nullptr});
}

template <bool TEMPORAL_STABILIZATION, int SIGNAL, int MODE, bool PROBE = false>
__global__ void __launch_bounds__(BLOCK_W* BLOCK_H, SPATIAL_MIN_BLOCKS) reblurPostBlurKernel(const __grid_constant__ ReblurConstants cb, const __grid_constant__ PostBlurParams p, int flags, int ctaY0) {
    pdlEntry();
    const bool quads = (flags & 1) != 0, robust = (flags & 2) != 0;
    Center s;
    const int2 cta = ctaTile<4>(ctaY0);
    s.px = cta.x * BLOCK_W + threadIdx.x;
    s.py = cta.y * BLOCK_H + threadIdx.y;

    bool skyTile = p.tiles.load(s.px >> 4, s.py >> 4) != 0.0f;
    s.data1 = loadData1<SIGNAL>(p, s.px, s.py);
    s.viewZ = unpackViewZ(cb, p.viewZ.load(s.px, s.py));
    s.nonLinearAccumSpeed = quadSmoothedAccumSpeed(cb, s.data1, s.viewZ, quads);
    if (skyTile || !inDenoisingRange(cb, s.viewZ) || s.px > cb.rectSizeMinusOne[0] || s.py > cb.rectSizeMinusOne[1]) return;

    setupCenter(cb, s, p.normalRoughness, cb.rotatorPost);

    p.outNormalRoughness.storeRaw(s.px, s.py, p.normalRoughness.loadRaw(s.px, s.py));
    if (!TEMPORAL_STABILIZATION) p.outInternalData.store(s.px, s.py, packInternalData(cb, s.data1.x, s.data1.y, s.materialID));

    if constexpr ((SIGNAL & SIGNAL_DIFF) != 0) spatialFilter<POST_BLUR, DIFF, false, MODE, PROBE>(cb, s, p.geom, p.normalRoughness, p.inDiff, p.outDiff, nullptr, &p.outDiffCopy, TEMPORAL_STABILIZATION, robust, nullptr, ShIo{&p.inDiffSh, &p.outDiffSh, &p.outDiffShCopy});
    if constexpr ((SIGNAL & SIGNAL_SPEC) != 0) spatialFilter<POST_BLUR, SPEC, false, MODE, PROBE>(cb, s, p.geom, p.normalRoughness, p.inSpec, p.outSpec, nullptr, &p.outSpecCopy, TEMPORAL_STABILIZATION, robust, nullptr, ShIo{&p.inSpecSh, &p.outSpecSh, &p.outSpecShCopy});
}

// REBLUR_SplitScreen.cs.hlsl:21-56: the noisy input (range-masked) left of CommonSettings::splitScreen
__global__ void __launch_bounds__(BLOCK_W* BLOCK_H) reblurSplitScreenKernel(const __grid_constant__ ReblurConstants cb, const __grid_constant__ SplitScreenParams p, int signal, int ctaY0) {
    pdlEntry();
    const int px = blockIdx.x * BLOCK_W + threadIdx.x, py = (ctaY0 + blockIdx.y) * BLOCK_H + threadIdx.y;
    const float u = ((float)px + 0.5f) * cb.rectSizeInv[0];
    if (u > cb.splitScreen || px > cb.rectSizeMinusOne[0] || py > cb.rectSizeMinusOne[1]) return;
    const float inRange = inDenoisingRange(cb, unpackViewZ(cb, p.viewZ.load(px, py))) ? 1.0f : 0.0f;
    // the RADIANCE permutation serves the occlusion denoisers too ( Reblur_DiffuseSpecularOcclusion.hpp:215-226 ): their textures are single-channel or SNORM,
    // a graphics API converts on access, here the views carry the format
    auto copy = [&](const TexRGBA16F& src, const TexRGBA16F& dst, uint32_t checkerboard) {
        const int sx = px >> (checkerboard != 2u ? 1 : 0);
        if (src.fmt == (uint32_t)nrd::Format::RGBA16_SFLOAT && dst.fmt == (uint32_t)nrd::Format::RGBA16_SFLOAT) dst.store(px, py, src.load(sx, py) * inRange);
        else anyStore4(dst, px, py, (src.inside(sx, py) ? anyFetch4(src, sx, py) : f4(0.0f)) * inRange);
    };
    if (signal & SIGNAL_DIFF) copy(p.inDiff, p.outDiff, cb.diffCheckerboard);
    if (signal & SIGNAL_SPEC) copy(p.inSpec, p.outSpec, cb.specCheckerboard);
    // NRD_MODE = SH ( :44-52 ): the SH1 inputs, same addressing
    if ((signal & SIGNAL_DIFF) && p.inDiffSh.data) p.outDiffSh.store(px, py, p.inDiffSh.load(px >> (cb.diffCheckerboard != 2u ? 1 : 0), py) * inRange);
    if ((signal & SIGNAL_SPEC) && p.inSpecSh.data) p.outSpecSh.store(px, py, p.inSpecSh.load(px >> (cb.specCheckerboard != 2u ? 1 : 0), py) * inRange);
}

// Reads ( and optionally clears ) g_mirrorProbe: out[0] = taps, out[1] = taps that took the "mirrored" branch, then the same pair per ( pass, lobe )
// slot at out[2 + 2 * slot] ( 14 values; slot = pass * 2 + lobe, pass 0 pre-pass / 1 blur / 2 post-blur, lobe 0 diffuse / 1 specular )
bool readMirrorProbe(unsigned long long* out, bool reset) {
    unsigned long long v[6][2], zero[6][2] = {};
    if (cudaMemcpyFromSymbol(v, g_mirrorProbe, sizeof(v)) != cudaSuccess) return false;
    if (out) {
        out[0] = out[1] = 0ull;
        for (int k = 0; k < 6; k++) {
            out[0] += v[k][0];
            out[1] += v[k][1];
            out[2 + 2 * k] = v[k][0];
            out[3 + 2 * k] = v[k][1];
        }
    }
    return !reset || cudaMemcpyToSymbol(g_mirrorProbe, zero, sizeof(zero)) == cudaSuccess;
}

// ---------------------------------------------------------------------------------------------------------------
// Host launchers (called by the executor)
// ---------------------------------------------------------------------------------------------------------------
void launchReblurClassifyTiles(const ReblurConstants& cb, const ClassifyTilesParams& p, Rows rows, cudaStream_t stream) {
    const RowGrid g = rowGrid(rows, cb.rectSizeMinusOne[1] + 1, 16);
    if (!g.count) return;
    const int tilesW = (cb.rectSizeMinusOne[0] + 16) / 16;
    launchK(reblurClassifyTilesKernel, dim3((tilesW + 7) / 8, g.count), 256, 0, stream, cb, p, g.ctaY0, tilesW);
}
// rows [ row0, row1 ) of the plane ( clamped to the rect )
void launchReblurGeometryPlane(const ReblurConstants& cb, const GeometryPlaneParams& p, int row0, int row1, cudaStream_t stream) {
    row0 = row0 < 0 ? 0 : row0;
    row1 = row1 > cb.rectSizeMinusOne[1] + 1 ? cb.rectSizeMinusOne[1] + 1 : row1;
    if (row1 <= row0) return;
    launchK(reblurGeometryPlaneKernel, dim3((cb.rectSizeMinusOne[0] + 32) / 32, (row1 - row0 + 7) / 8), 256, 0, stream, cb, p, row0, row1);
}
void launchReblurSplitScreen(const ReblurConstants& cb, const SplitScreenParams& p, int signal, Rows rows, cudaStream_t stream) {
    const RowGrid g = rowGrid(rows, cb.rectSizeMinusOne[1] + 1, BLOCK_H);
    if (!g.count) return;
    launchK(reblurSplitScreenKernel, dim3((cb.rectSizeMinusOne[0] + BLOCK_W) / BLOCK_W, g.count), dim3(BLOCK_W, BLOCK_H), 0, stream, cb, p, signal, g.ctaY0);
}
// `flags`: bit 0 quad smoothing, bit 1 robust mirror test, bit 2 mirror probe, bits 4-5 NRD_MODE of the permutation ( MODE_* )
void launchReblurPrePass(const ReblurConstants& cb, const PrePassParams& p, int signal, int flags, Rows rows, cudaStream_t stream) {
    const RowGrid g = rowGrid(rows, cb.rectSizeMinusOne[1] + 1, BLOCK_H);
    if (!g.count) return;
    const dim3 grid((cb.rectSizeMinusOne[0] + BLOCK_W) / BLOCK_W, g.count), block(BLOCK_W, BLOCK_H);
    const bool cbOn = cb.diffCheckerboard != 2u || cb.specCheckerboard != 2u;  // CheckerboardMode::BLACK / WHITE set both (Reblur.cpp:301-313)
    const int mode = (flags >> 4) & 3;
    if ((flags & 4) && signal == SIGNAL_BOTH && mode == MODE_RADIANCE && !cbOn) {  // NRDCU_FLAG_PROBE_MIRROR
        launchK(reblurPrePassKernel<false, SIGNAL_BOTH, MODE_RADIANCE, true>, grid, block, 0, stream, cb, p, flags, g.ctaY0);
        return;
    }
    if (mode == MODE_DO) {   // REBLUR_DIFFUSE_DIRECTIONAL_OCCLUSION: the only denoiser of this mode has one lobe ( the occlusion denoisers have no pre-pass )
        if (cbOn) launchK(reblurPrePassKernel<true, SIGNAL_DIFF, MODE_DO>, grid, block, 0, stream, cb, p, flags, g.ctaY0);
        else launchK(reblurPrePassKernel<false, SIGNAL_DIFF, MODE_DO>, grid, block, 0, stream, cb, p, flags, g.ctaY0);
        return;
    }
    withSignal(signal, [&](auto sig) {
        constexpr int S = decltype(sig)::value;
        if (mode == MODE_SH) {
            if (cbOn) launchK(reblurPrePassKernel<true, S, MODE_SH>, grid, block, 0, stream, cb, p, flags, g.ctaY0);
            else launchK(reblurPrePassKernel<false, S, MODE_SH>, grid, block, 0, stream, cb, p, flags, g.ctaY0);
        } else {
            if (cbOn) launchK(reblurPrePassKernel<true, S, MODE_RADIANCE>, grid, block, 0, stream, cb, p, flags, g.ctaY0);
            else launchK(reblurPrePassKernel<false, S, MODE_RADIANCE>, grid, block, 0, stream, cb, p, flags, g.ctaY0);
        }
    });
}
void launchReblurBlur(const ReblurConstants& cb, const BlurParams& p, int signal, int flags, Rows rows, cudaStream_t stream) {
    // the grid covers the rect rounded up to the reference's 16-row groups: their threads below the rect still copy viewZ
    const RowGrid g = rowGrid(rows, (cb.rectSizeMinusOne[1] + 16) & ~15, BLOCK_H);
    if (!g.count) return;
    const int mode = (flags >> 4) & 3;
    const dim3 grid((cb.rectSizeMinusOne[0] + BLOCK_W) / BLOCK_W, g.count), block(BLOCK_W, BLOCK_H);
    if ((flags & 4) && signal == SIGNAL_BOTH && mode == MODE_RADIANCE) {
        launchK(reblurBlurKernel<SIGNAL_BOTH, MODE_RADIANCE, true>, grid, block, 0, stream, cb, p, flags, g.ctaY0);
        return;
    }
    if (mode == MODE_DO) {
        launchK(reblurBlurKernel<SIGNAL_DIFF, MODE_DO>, grid, block, 0, stream, cb, p, flags, g.ctaY0);
        return;
    }
    withSignal(signal, [&](auto sig) {
        constexpr int S = decltype(sig)::value;
        if (mode == MODE_SH) launchK(reblurBlurKernel<S, MODE_SH>, grid, block, 0, stream, cb, p, flags, g.ctaY0);
        else if (mode == MODE_OCCLUSION) launchK(reblurBlurKernel<S, MODE_OCCLUSION>, grid, block, 0, stream, cb, p, flags, g.ctaY0);
        else launchK(reblurBlurKernel<S, MODE_RADIANCE>, grid, block, 0, stream, cb, p, flags, g.ctaY0);
    });
}
void launchReblurPostBlur(const ReblurConstants& cb, const PostBlurParams& p, int signal, bool temporalStabilization, int flags, Rows rows, cudaStream_t stream) {
    const RowGrid g = rowGrid(rows, cb.rectSizeMinusOne[1] + 1, BLOCK_H);
    if (!g.count) return;
    const int mode = (flags >> 4) & 3;
    const dim3 grid((cb.rectSizeMinusOne[0] + BLOCK_W) / BLOCK_W, g.count), block(BLOCK_W, BLOCK_H);
    if ((flags & 4) && signal == SIGNAL_BOTH && mode == MODE_RADIANCE && temporalStabilization) {
        launchK(reblurPostBlurKernel<true, SIGNAL_BOTH, MODE_RADIANCE, true>, grid, block, 0, stream, cb, p, flags, g.ctaY0);
        return;
    }
    if (mode == MODE_DO) {
        if (temporalStabilization) launchK(reblurPostBlurKernel<true, SIGNAL_DIFF, MODE_DO>, grid, block, 0, stream, cb, p, flags, g.ctaY0);
        else launchK(reblurPostBlurKernel<false, SIGNAL_DIFF, MODE_DO>, grid, block, 0, stream, cb, p, flags, g.ctaY0);
        return;
    }
    withSignal(signal, [&](auto sig) {
        constexpr int S = decltype(sig)::value;
        if (mode == MODE_OCCLUSION) launchK(reblurPostBlurKernel<false, S, MODE_OCCLUSION>, grid, block, 0, stream, cb, p, flags, g.ctaY0);   // the occlusion graph has no stabilization pass
        else if (mode == MODE_SH) {
            if (temporalStabilization) launchK(reblurPostBlurKernel<true, S, MODE_SH>, grid, block, 0, stream, cb, p, flags, g.ctaY0);
            else launchK(reblurPostBlurKernel<false, S, MODE_SH>, grid, block, 0, stream, cb, p, flags, g.ctaY0);
        } else {
            if (temporalStabilization) launchK(reblurPostBlurKernel<true, S, MODE_RADIANCE>, grid, block, 0, stream, cb, p, flags, g.ctaY0);
            else launchK(reblurPostBlurKernel<false, S, MODE_RADIANCE>, grid, block, 0, stream, cb, p, flags, g.ctaY0);
        }
    });
}

}  // namespace nrdk
