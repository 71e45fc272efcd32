// REBLUR temporal accumulation on sm_100a (NRD_MODE=RADIANCE; NRD_SIGNAL = DIFF / SPEC / BOTH is a template parameter).
//
// Replaces External/NRD/Shaders/REBLUR_TemporalAccumulation.cs.hlsl:68-995: surface-motion and virtual (specular)
// motion reprojection with per-tap occlusion tests on the previous frame's viewZ / normals / material IDs, thin-lens
// curvature estimation, confidence-driven history length, 12-tap Catmull-Rom history fetch with bilinear fallback,
// firefly suppression and the fast (short) history.
//
// Mapping to the GPU: one thread per pixel, 32x8 CTAs. The 3x3 neighbourhood of { normal, hit distance for tracking }
// the reference stages in group-shared memory is staged here in shared memory too (34x10 float4 tile, decoded once per
// texel); everything else is per-pixel scattered history reads that land in L2 (previous-frame planes at pixel + MV).
// This kernel is ALU-bound (atan/log/pow/exp2 heavy), not bandwidth-bound — see DESIGN.md.
#include "reblur_common.cuh"

namespace nrdk {

namespace {

constexpr int BLOCK_W = 32, BLOCK_H = 8, BORDER = 1;
constexpr int TILE_W = BLOCK_W + 2 * BORDER, TILE_H = BLOCK_H + 2 * BORDER;

// Requires rng.init( pixelPos, frameIndex ): picks ONE texel of the bilinear footprint (REBLUR_USE_STF = 1). The reference turns ( origin + 0.5 ) / filterSize
// back into a texel through a nearest sampler ( clamp addressing ) of a texture that is `texW x texH` texels ( the RESOURCE size: with dynamic resolution
// the rect the filter works in is smaller ), after scaling the uv by `scale` ( gResolutionScalePrev for the previous frame, TA:665; 1 for the current one, TA:440 )
NRD_DEV void stochasticBilinearTexel(Rng& rng, float2 uv, float2 filterSize, float2 scale, int texW, int texH, int& tx, int& ty) {
    Bilinear f = getBilinearFilter(uv, filterSize);
    float r0 = rng.next();
    float r1 = rng.next();
    float ox = f.origin.x + step(r0, f.weights.x), oy = f.origin.y + step(r1, f.weights.y);
    float2 uvq = make_float2((ox + 0.5f) / filterSize.x, (oy + 0.5f) / filterSize.y) * scale;
    tx = (int)floorf(uvq.x * (float)texW);
    ty = (int)floorf(uvq.y * (float)texH);
}

NRD_DEV float4 gather4(const TexR32F& t, int x0, int y0) {
    return make_float4(t.fetchClamped(x0, y0), t.fetchClamped(x0 + 1, y0), t.fetchClamped(x0, y0 + 1), t.fetchClamped(x0 + 1, y0 + 1));
}
NRD_DEV uint4 gather4(const TexR16U& t, int x0, int y0) {
    return make_uint4(t.fetchClamped(x0, y0), t.fetchClamped(x0 + 1, y0), t.fetchClamped(x0, y0 + 1), t.fetchClamped(x0 + 1, y0 + 1));
}

}  // namespace

#ifndef TA_MIN_BLOCKS
#    define TA_MIN_BLOCKS 4  // 64 regs (4 CTAs / SM): 383 us vs 401 us at 80 regs (3 CTAs) and 467 us at 128 regs (2 CTAs) for a 1440p frame on B200
#endif
// OPTIONAL: checkerboard resolve speed-up and the application's guide textures (confidence, threshold mix); compiled out of the plain kernel
// SH ( NRD_MODE = SH ): the lobe's second RGBA16F accumulates with the lobe's speed from a bilinear ( custom weights ) history fetch and follows the
// firefly clamp of the luma ( REBLUR_TemporalAccumulation.cs.hlsl:781-783, 801-804, 819-821, 831-833, 923-925, 940-943, 957-959, 968-970 )
// MODE ( NRD_MODE ): OCCLUSION / DO read and write their signal and fast history through the format-polymorphic Sig / FastSig, skip the firefly suppressor
// ( NRD_SUPPORTS_ANTIFIREFLY = 0, REBLUR_Config.hlsli:38-41 ); OCCLUSION additionally has no pre-pass in front of it — it reads the ( possibly half-width,
// checkerboarded ) input itself and resolves the missing pixels from their row neighbours ( TA:331-345, 359-378, 893-912 ) — and writes no data2.
template <bool OPTIONAL, int SIGNAL, int MODE>
__global__ void __launch_bounds__(BLOCK_W* BLOCK_H, TA_MIN_BLOCKS) reblurTemporalAccumulationKernel(const __grid_constant__ ReblurConstants cb,
                                                                                      const __grid_constant__ TemporalAccumulationParams p, int ctaY0) {
    pdlEntry();
    __shared__ float4 sNormalHitDist[TILE_H][TILE_W];
    constexpr bool SH = MODE == MODE_SH, FIXED = Sig<MODE>::FIXED, OCC = MODE == MODE_OCCLUSION;

    const int2 cta = ctaTile<1>(ctaY0);
    const int px = cta.x * BLOCK_W + threadIdx.x, py = cta.y * BLOCK_H + threadIdx.y;
    const float2 rectSize = make_float2(cb.rectSize[0], cb.rectSize[1]);
    const float2 rectSizeInv = make_float2(cb.rectSizeInv[0], cb.rectSizeInv[1]);
    const float2 rectSizePrev = make_float2(cb.rectSizePrev[0], cb.rectSizePrev[1]);
    const float2 resourceSizeInvPrev = make_float2(cb.resourceSizeInvPrev[0], cb.resourceSizeInvPrev[1]);
    const float2 resolutionScalePrev = make_float2(cb.resolutionScalePrev[0], cb.resolutionScalePrev[1]);
    const float3 cameraDelta = make_float3(cb.cameraDelta[0], cb.cameraDelta[1], cb.cameraDelta[2]);

    // ---- Preload (TA:38-66): { N, hit distance for tracking or INF } of the clamped 34x10 neighbourhood ----
    {
        const int baseX = cta.x * BLOCK_W - BORDER, baseY = cta.y * BLOCK_H - BORDER;
        const int tid = threadIdx.y * BLOCK_W + threadIdx.x;
        for (int i = tid; i < TILE_W * TILE_H; i += BLOCK_W * BLOCK_H) {
            int sx = i % TILE_W, sy = i / TILE_W;
            int gx = clampi(baseX + sx, 0, cb.rectSizeMinusOne[0]), gy = clampi(baseY + sy, 0, cb.rectSizeMinusOne[1]);
            float3 N = xyz(unpackNormalRoughness(p.normalRoughness.loadRaw(gx, gy)));
            if constexpr ((SIGNAL & SIGNAL_SPEC) != 0) {
                float hitDist;
                if constexpr (OCC) hitDist = Sig<MODE>::load(p.inSpec, gx >> (cb.specCheckerboard != 2u ? 1 : 0), gy).w;   // TA:46-55
                else hitDist = cb.specPrepassBlurRadius == 0.0f ? Sig<MODE>::load(p.inSpec, gx, gy).w : p.inSpecHitDistForTracking.load(gx, gy);
                float z = unpackViewZ(cb, p.viewZ.load(gx, gy));
                sNormalHitDist[sy][sx] = f4(N, (hitDist == 0.0f || !inDenoisingRange(cb, z)) ? NRD_INF : hitDist);
            } else
                sNormalHitDist[sy][sx] = f4(N, 0.0f);  // the tracking distance is a specular-only quantity (TA:43-63)
        }
    }
    __syncthreads();

    if (p.tiles.load(px >> 4, py >> 4) != 0.0f || px > cb.rectSizeMinusOne[0] || py > cb.rectSizeMinusOne[1]) return;
    const float viewZ = unpackViewZ(cb, p.viewZ.load(px, py));
    if (!inDenoisingRange(cb, viewZ)) return;

    // Current position
    const float2 pixelUv = make_float2(px + 0.5f, py + 0.5f) * rectSizeInv;
    const float3 Xv = reconstructViewPosition(pixelUv, cb.frustum, viewZ, cb.orthoMode);
    const float3 X = rotate(cb.viewToWorld, Xv);

    // 3x3 min hit distance, 2x2 averaged (unnormalised) normal
    float3 Navg = f3(0.0f);
    float hitDistForTracking = NRD_INF;
#pragma unroll
    for (int j = 0; j <= 2; j++)
#pragma unroll
        for (int i = 0; i <= 2; i++) {
            float4 d = sNormalHitDist[threadIdx.y + j][threadIdx.x + i];
            if (i < 2 && j < 2) Navg += xyz(d) * 0.25f;
            hitDistForTracking = fminf(hitDistForTracking, d.w);
        }

    float materialID;
    const float4 normalAndRoughness = unpackNormalRoughness(p.normalRoughness.loadRaw(px, py), materialID);
    const float3 N = xyz(normalAndRoughness);
    const float roughness = normalAndRoughness.w;
    const float roughnessModified = modifiedRoughnessFromNormalVariance(roughness, Navg);

    Rng rng;
    rng.init((uint32_t)px, (uint32_t)py, cb.frameIndex);

    hitDistForTracking = hitDistForTracking == NRD_INF ? 0.0f : hitDistForTracking;
    const float hitDistNormalization = hitDistanceNormalization(viewZ, cb.hitDistSettings, roughness);
    hitDistForTracking *= (OCC || cb.specPrepassBlurRadius == 0.0f) ? hitDistNormalization : 1.0f;   // TA:135-139
    if constexpr ((SIGNAL & SIGNAL_SPEC) != 0) p.outSpecHitDistForTracking.store(px, py, hitDistForTracking);

    // Previous position and surface motion uv
    float4 mvRaw = p.mv.load(px, py);
    float3 mv = make_float3(mvRaw.x * cb.mvScale[0], mvRaw.y * cb.mvScale[1], mvRaw.z * cb.mvScale[2]);
    float3 Xprev = X;
    float2 smbPixelUv = pixelUv + xy(mv);
    if (cb.mvScale[3] == 0.0f) {
        if (cb.mvScale[2] == 0.0f) mv.z = affine(cb.worldToViewPrev, X).z - viewZ;
        float viewZprev = viewZ + mv.z;
        float3 Xvprevlocal = reconstructViewPosition(smbPixelUv, cb.frustumPrev, viewZprev, cb.orthoMode);
        Xprev = rotateInverse(cb.worldToViewPrev, Xvprevlocal) + cameraDelta;
    } else {
        Xprev += mv;
        smbPixelUv = screenUv(cb.worldToClipPrev, Xprev);
    }

    // Previous viewZ over the 4x4 Catmull-Rom footprint = four 2x2 gathers (TA:165-194); k = bilinear origin
    const Bilinear smbBilinearFilter = getBilinearFilter(smbPixelUv, rectSizePrev);
    const int kx = (int)floorf(smbPixelUv.x * rectSizePrev.x - 0.5f), ky = (int)floorf(smbPixelUv.y * rectSizePrev.y - 0.5f);
    const float4 z0 = gather4(p.prevViewZ, kx - 1, ky - 1), z1 = gather4(p.prevViewZ, kx + 1, ky - 1);
    const float4 z2 = gather4(p.prevViewZ, kx - 1, ky + 1), z3 = gather4(p.prevViewZ, kx + 1, ky + 1);
    const float3 prevViewZ0 = make_float3(unpackViewZ(cb, z0.y), unpackViewZ(cb, z0.z), unpackViewZ(cb, z0.w));
    const float3 prevViewZ1 = make_float3(unpackViewZ(cb, z1.x), unpackViewZ(cb, z1.z), unpackViewZ(cb, z1.w));
    const float3 prevViewZ2 = make_float3(unpackViewZ(cb, z2.x), unpackViewZ(cb, z2.y), unpackViewZ(cb, z2.w));
    const float3 prevViewZ3 = make_float3(unpackViewZ(cb, z3.x), unpackViewZ(cb, z3.y), unpackViewZ(cb, z3.z));

    // Previous normals of the 2x2 bilinear footprint against Navg
    float smbNoN;
    float4 smbNoN2x2;
    {
        int bx = (int)smbBilinearFilter.origin.x, by = (int)smbBilinearFilter.origin.y;
        float3 n00 = xyz(unpackNormalRoughness(p.prevNormalRoughness.loadRaw(bx, by)));
        float3 n10 = xyz(unpackNormalRoughness(p.prevNormalRoughness.loadRaw(bx + 1, by)));
        float3 n01 = xyz(unpackNormalRoughness(p.prevNormalRoughness.loadRaw(bx, by + 1)));
        float3 n11 = xyz(unpackNormalRoughness(p.prevNormalRoughness.loadRaw(bx + 1, by + 1)));
        smbNoN2x2 = make_float4(dot(n00, Navg), dot(n10, Navg), dot(n01, Navg), dot(n11, Navg));
        smbNoN = applyBilinear(smbNoN2x2.x, smbNoN2x2.y, smbNoN2x2.z, smbNoN2x2.w, smbBilinearFilter);
    }

    // Parallax
    const float smbParallaxInPixels1 = parallaxInPixels(Xprev + cameraDelta, cb.orthoMode == 0.0f ? smbPixelUv : pixelUv, cb.worldToClipPrev, rectSize);
    const float smbParallaxInPixels2 = parallaxInPixels(Xprev - cameraDelta, cb.orthoMode == 0.0f ? pixelUv : smbPixelUv, cb.worldToClip, rectSize);
    const float smbParallaxInPixelsMax = fmaxf(smbParallaxInPixels1, smbParallaxInPixels2);
    const float smbParallaxInPixelsMin = fminf(smbParallaxInPixels1, smbParallaxInPixels2);

    // Disocclusion threshold
    const float pixelSize = pixelRadiusToWorld(cb.unproject, cb.orthoMode, 1.0f, viewZ);
    const float frustumSize = frustumSizeAt(cb.minRectDimMulUnproject, cb.orthoMode, viewZ);

    float disocclusionThresholdMix = 0.0f;
    if (materialID == cb.strandMaterialID) disocclusionThresholdMix = saturate(0.5f * pixelSize / (cb.strandThickness + NRD_EPS));
    if (OPTIONAL && cb.hasDisocclusionThresholdMix) disocclusionThresholdMix = p.disocclusionThresholdMix.load(px, py);
    float disocclusionThreshold = lerp(cb.disocclusionThreshold, cb.disocclusionThresholdAlternate, disocclusionThresholdMix);
    if (materialID == cb.strandMaterialID) disocclusionThreshold = lerp(0.25f, disocclusionThreshold, smoothStep01(smbParallaxInPixelsMax));

    const float smallParallax = linearStep(0.25f, 0.0f, smbParallaxInPixelsMax);
    const float cosMaxAngle = cosf(degToRad(89.0f)) - 0.25f * smallParallax;

    const float3 V = viewVector(cb, X);
    const float NoV = fabsf(dot(N, V));
    const float NoVstrict = lerp(NoV, 1.0f, saturate(smbParallaxInPixelsMax / 30.0f));

    // Disocclusion
    float4 smbDisocclusionThreshold = make_float4(smbNoN2x2.x > cosMaxAngle, smbNoN2x2.y > cosMaxAngle, smbNoN2x2.z > cosMaxAngle, smbNoN2x2.w > cosMaxAngle);
    smbDisocclusionThreshold *= isInScreenBilinear(smbBilinearFilter.origin, rectSizePrev);
    smbDisocclusionThreshold *= disocclusionThresholdAt(disocclusionThreshold, frustumSize, NoVstrict);
    smbDisocclusionThreshold -= NRD_EPS;

    const float3 Xvprev = affine(cb.worldToViewPrev, Xprev);
    auto occl3 = [&](float3 z, float thr) {
        return make_float3((fabsf(z.x - Xvprev.z) <= thr && inDenoisingRange(cb, z.x)) ? 1.0f : 0.0f, (fabsf(z.y - Xvprev.z) <= thr && inDenoisingRange(cb, z.y)) ? 1.0f : 0.0f,
                           (fabsf(z.z - Xvprev.z) <= thr && inDenoisingRange(cb, z.z)) ? 1.0f : 0.0f);
    };
    float3 smbOcclusion0 = occl3(prevViewZ0, smbDisocclusionThreshold.x);
    float3 smbOcclusion1 = occl3(prevViewZ1, smbDisocclusionThreshold.y);
    float3 smbOcclusion2 = occl3(prevViewZ2, smbDisocclusionThreshold.z);
    float3 smbOcclusion3 = occl3(prevViewZ3, smbDisocclusionThreshold.w);

    // Disocclusion: material ID
    const uint4 id0 = gather4(p.prevInternalData, kx - 1, ky - 1), id1 = gather4(p.prevInternalData, kx + 1, ky - 1);
    const uint4 id2 = gather4(p.prevInternalData, kx - 1, ky + 1), id3 = gather4(p.prevInternalData, kx + 1, ky + 1);
    {
        const float minMaterialID = fminf(cb.specMinMaterial, cb.diffMinMaterial);
        auto m = [&](uint32_t v) { return compareMaterials(materialID, (float)((v >> 12) & 15u) * (1.0f / 15.0f) * 15.0f, minMaterialID) ? 1.0f : 0.0f; };
        smbOcclusion0 *= make_float3(m(id0.y), m(id0.z), m(id0.w));
        smbOcclusion1 *= make_float3(m(id1.x), m(id1.z), m(id1.w));
        smbOcclusion2 *= make_float3(m(id2.x), m(id2.y), m(id2.w));
        smbOcclusion3 *= make_float3(m(id3.x), m(id3.y), m(id3.z));
    }

    const float4 smbOcclusionWeights = bilinearCustomWeights(smbBilinearFilter, make_float4(smbOcclusion0.z, smbOcclusion1.y, smbOcclusion2.y, smbOcclusion3.x));
    const float3 osum = smbOcclusion0 + smbOcclusion1 + smbOcclusion2 + smbOcclusion3;
    const bool smbAllowCatRom = (osum.x + osum.y + osum.z) > 11.5f;

    float fbits = smbOcclusion0.z * 1.0f;
    fbits += smbOcclusion1.y * 2.0f;
    fbits += smbOcclusion2.y * 4.0f;
    fbits += smbOcclusion3.x * 8.0f;

    // Accumulation speed
    const float3 internalData00 = unpackInternalData(id0.w), internalData10 = unpackInternalData(id1.z);
    const float3 internalData01 = unpackInternalData(id2.y), internalData11 = unpackInternalData(id3.x);
    float diffAccumSpeed = applyCustomWeights(internalData00.x, internalData10.x, internalData01.x, internalData11.x, smbOcclusionWeights);
    float smbSpecAccumSpeed = applyCustomWeights(internalData00.y, internalData10.y, internalData01.y, internalData11.y, smbOcclusionWeights);

    // Footprint quality
    const float3 smbVprev = viewVectorPrev(cb, Xprev, cameraDelta);
    const float NoVprev = fabsf(dot(N, smbVprev));
    float sizeQuality = (NoVprev + 1e-3f) / (NoV + 1e-3f);
    sizeQuality *= sizeQuality;
    sizeQuality = lerp(0.1f, 1.0f, saturate(sizeQuality));
    float smbFootprintQuality = applyBilinear(smbOcclusion0.z, smbOcclusion1.y, smbOcclusion2.y, smbOcclusion3.x, smbBilinearFilter);
    smbFootprintQuality = sqrt01(smbFootprintQuality);
    smbFootprintQuality *= sizeQuality;

    // Checkerboard resolve of the occlusion mode ( TA:331-345 ): the traced row neighbours and their disocclusion weights
    int cbX0 = 0, cbX1 = 0;
    float2 wc = f2(0.0f);
    if constexpr (OCC) {
        const int x0 = max(px - 1, 0), x1 = min(px + 1, cb.rectSizeMinusOne[0]);
        const float viewZ0 = unpackViewZ(cb, p.viewZ.load(x0, py)), viewZ1 = unpackViewZ(cb, p.viewZ.load(x1, py));
        const float threshold = disocclusionThresholdAt(0.02f, frustumSize, NoV);  // NRD_DISOCCLUSION_THRESHOLD
        wc = make_float2(threshold >= fabsf(viewZ0 - viewZ) ? 1.0f : 0.0f, threshold >= fabsf(viewZ1 - viewZ) ? 1.0f : 0.0f);
        if (!inDenoisingRange(cb, viewZ0) || px < 1) wc.x = 0.0f;
        if (!inDenoisingRange(cb, viewZ1) || px >= cb.rectSizeMinusOne[0]) wc.y = 0.0f;
        wc = wc * positiveRcp(wc.x + wc.y);
        cbX0 = x0 >> 1;
        cbX1 = x1 >> 1;
    }
    // the lobe's input of this pixel: full width after a pre-pass, half width and resolved here in the occlusion mode
    auto currentSignal = [&](const TexRGBA16F& tex, uint32_t checkerboard, bool hasData) {
        if constexpr (!OCC) return Sig<MODE>::load(tex, px, py);
        else {
            float4 v = Sig<MODE>::load(tex, px >> (checkerboard != 2u ? 1 : 0), py);
            if (!hasData) {
                float4 s0 = Sig<MODE>::load(tex, cbX0, py), s1 = Sig<MODE>::load(tex, cbX1, py);
                if (wc.x == 0.0f) s0 = f4(0.0f);
                if (wc.y == 0.0f) s1 = f4(0.0f);
                v = s0 * wc.x + s1 * wc.y;
            }
            return v;
        }
    };

    // =============================================================================================== Specular
    float specAccumSpeedCorrected = 0.0f, curvature = 0.0f, virtualHistoryAmount = 0.0f;  // what a diffuse-only denoiser packs (TA:869-873)
    if constexpr ((SIGNAL & SIGNAL_SPEC) != 0) {
        float smbSpecHistoryConfidence = smbFootprintQuality;
        if (OPTIONAL && cb.hasHistoryConfidence) smbSpecHistoryConfidence = fminf(smbSpecHistoryConfidence, saturate(p.specConfidence.sampleLinear(smbPixelUv)));
        smbSpecAccumSpeed *= lerp(smbSpecHistoryConfidence, 1.0f, 1.0f / (1.0f + smbSpecAccumSpeed));

        // Checkerboard ( RADIANCE mode: the pre-pass has already resolved the half-width input, only the accumulation speed changes; TA:329-357 )
        const bool specHasData = !OPTIONAL || cb.specCheckerboard == 2u || (((uint32_t)(px ^ py) ^ cb.frameIndex) & 1u) == cb.specCheckerboard;
        const float4 spec = currentSignal(p.inSpec, cb.specCheckerboard, specHasData);

        // Curvature estimation along predicted motion (TA:387-467)
        curvature = 0.0f;
        {
            float2 uvForZeroParallax = cb.orthoMode == 0.0f ? smbPixelUv : pixelUv;
            float2 deltaUv = uvForZeroParallax - screenUv(cb.worldToClipPrev, Xprev + cameraDelta);
            deltaUv *= rectSize;
            deltaUv /= fmaxf(smbParallaxInPixels1, 1.0f / 256.0f);

            float3 n10, x10, n01, x01;
            {
                float3 xv = reconstructViewPosition(pixelUv + make_float2(1, 0) * rectSizeInv, cb.frustum, 1.0f, cb.orthoMode);
                float3 x = rotate(cb.viewToWorld, xv);
                float3 v = viewVector(cb, x);
                float3 o = cb.orthoMode == 0.0f ? f3(0.0f) : x;
                x10 = o + v * dot(X - o, N) / dot(N, v);
                n10 = xyz(sNormalHitDist[threadIdx.y + BORDER][threadIdx.x + BORDER + 1]);
            }
            {
                float3 xv = reconstructViewPosition(pixelUv + make_float2(0, 1) * rectSizeInv, cb.frustum, 1.0f, cb.orthoMode);
                float3 x = rotate(cb.viewToWorld, xv);
                float3 v = viewVector(cb, x);
                float3 o = cb.orthoMode == 0.0f ? f3(0.0f) : x;
                x01 = o + v * dot(X - o, N) / dot(N, v);
                n01 = xyz(sNormalHitDist[threadIdx.y + BORDER + 1][threadIdx.x + BORDER]);
            }

            float2 ww = fabs2(deltaUv) + 1.0f / 256.0f;
            ww /= ww.x + ww.y;

            float3 x = x10 * ww.x + x01 * ww.y;
            float3 n = normalize(n10 * ww.x + n01 * ww.y);

            float2 motionUvHigh = pixelUv + smbParallaxInPixelsMin * deltaUv * rectSizeInv;
            if (smbParallaxInPixelsMin > sqrtf(2.0f) && isInScreenNearest(motionUvHigh)) {
                float2 resolutionScale = make_float2(cb.resolutionScale[0], cb.resolutionScale[1]);
                float2 uvScaled = min2(motionUvHigh * resolutionScale, resolutionScale - 0.5f * make_float2(cb.resourceSizeInv[0], cb.resourceSizeInv[1]));

                float zHigh = unpackViewZ(cb, p.viewZ.sampleLinear(uvScaled));
                float3 xHigh = rotate(cb.viewToWorld, reconstructViewPosition(motionUvHigh, cb.frustum, zHigh, cb.orthoMode));

                int tx, ty;
                stochasticBilinearTexel(rng, uvScaled, rectSize, f2(1.0f), p.normalRoughness.w, p.normalRoughness.h, tx, ty);
                float3 nHigh = xyz(unpackNormalRoughness(p.normalRoughness.fetchRawClamped(tx, ty)));

                float2 gp = geometryWeightParams(0.04f, frustumSize, X, N);
                float w = applyGeometryWeightLast(cb, 1.0f, zHigh, dot(N, xHigh), gp);
                bool cmp = w > 0.5f;
                n = cmp ? nHigh : n;
                x = cmp ? xHigh : x;
            }

            float3 edge = x - X;
            curvature = dot(n - N, edge) * positiveRcp(dot(edge, edge));

            if (curvature < 0.0f) {
                float2 uv1 = screenUv(cb.worldToClipPrev, getXvirtual(hitDistForTracking, curvature, X, X, N, V, roughness));
                float2 uv2 = screenUv(cb.worldToClipPrev, X);
                float a = length((uv1 - uv2) * rectSize);
                curvature *= (a < 5.0f * smbParallaxInPixelsMax + rectSizeInv.x) ? 1.0f : 0.0f;
            }
        }

        // Virtual motion - coordinates
        const float3 Xvirtual = getXvirtual(hitDistForTracking, curvature, X, Xprev, N, V, roughness);
        float hitDistanceToLobeSpreadInPixels = 1.0f / pixelRadiusToWorld(cb.unproject, cb.orthoMode, 1.0f, length(Xvirtual));

        float2 vmbPixelUv = screenUv(cb.worldToClipPrev, Xvirtual);
        vmbPixelUv = materialID == cb.cameraAttachedReflectionMaterialID ? smbPixelUv : vmbPixelUv;

        float2 vmbDelta = vmbPixelUv - smbPixelUv;
        const float vmbPixelsTraveled = length(vmbDelta * rectSize);

        const Bilinear vmbBilinearFilter = getBilinearFilter(vmbPixelUv, rectSizePrev);
        const int vx = (int)vmbBilinearFilter.origin.x, vy = (int)vmbBilinearFilter.origin.y;

        // The four previous-frame normal+roughness texels of the vmb footprint serve the roughness, normal and occlusion tests
        const uint32_t nrRaw00 = p.prevNormalRoughness.fetchRawClamped(vx, vy), nrRaw10 = p.prevNormalRoughness.fetchRawClamped(vx + 1, vy);
        const uint32_t nrRaw01 = p.prevNormalRoughness.fetchRawClamped(vx, vy + 1), nrRaw11 = p.prevNormalRoughness.fetchRawClamped(vx + 1, vy + 1);

        // Virtual motion - confidence: roughness (gather = clamp addressing)
        float virtualHistoryConfidence;
        float4 roughnessWeights;
        {
            float2 rp = relaxedRoughnessWeightParams(roughness * roughness, cb.roughnessFraction, NRD_ROUGHNESS_SENSITIVITY * 0.3f);
            float4 r = make_float4(roughnessFromRaw(nrRaw00), roughnessFromRaw(nrRaw10), roughnessFromRaw(nrRaw01), roughnessFromRaw(nrRaw11));
            roughnessWeights = make_float4(nonExponentialWeight(r.x * r.x, rp.x, rp.y), nonExponentialWeight(r.y * r.y, rp.x, rp.y), nonExponentialWeight(r.z * r.z, rp.x, rp.y),
                                           nonExponentialWeight(r.w * r.w, rp.x, rp.y));
            roughnessWeights = lerp(f4(1.0f), roughnessWeights, smoothStep01(vmbPixelsTraveled));
            virtualHistoryConfidence = applyBilinear(roughnessWeights.x, roughnessWeights.y, roughnessWeights.z, roughnessWeights.w, vmbBilinearFilter);
        }

        // Load (not gather) semantics for the normals: out of bounds -> 0
        float4 vmbN;
        float4 vmbNoN2x2;
        float vmbNoN;
        {
            auto ld = [&](int x, int y, uint32_t raw) { return unpackNormalRoughness(p.prevNormalRoughness.inside(x, y) ? raw : 0u); };
            float4 n00 = ld(vx, vy, nrRaw00), n10 = ld(vx + 1, vy, nrRaw10), n01 = ld(vx, vy + 1, nrRaw01), n11 = ld(vx + 1, vy + 1, nrRaw11);
            vmbNoN2x2 = make_float4(dot(xyz(n00), N), dot(xyz(n10), N), dot(xyz(n01), N), dot(xyz(n11), N));
            vmbNoN = applyBilinear(vmbNoN2x2.x, vmbNoN2x2.y, vmbNoN2x2.z, vmbNoN2x2.w, vmbBilinearFilter);
            vmbN = applyBilinear(n00, n10, n01, n11, vmbBilinearFilter);
            float3 nn = xyz(vmbN);
            nn = nn * (1.0f / sqrtf(dot(nn, nn) + 1e-9f));
            vmbN = f4(nn, vmbN.w);
        }

        // Virtual motion - disocclusion
        float4 vmbOcclusionWeights;
        float vmbSpecAccumSpeed;
        bool vmbAllowCatRom;
        {
            float4 thr = make_float4(vmbNoN2x2.x > cosMaxAngle, vmbNoN2x2.y > cosMaxAngle, vmbNoN2x2.z > cosMaxAngle, vmbNoN2x2.w > cosMaxAngle);
            thr *= make_float4(step(0.5f, roughnessWeights.x), step(0.5f, roughnessWeights.y), step(0.5f, roughnessWeights.z), step(0.5f, roughnessWeights.w));
            thr *= isInScreenBilinear(vmbBilinearFilter.origin, rectSizePrev);
            thr *= disocclusionThreshold * frustumSize;
            thr *= lerp(0.1f, 1.0f, NoV);
            thr -= NRD_EPS;

            float4 zr = gather4(p.prevViewZ, vx, vy);
            float4 vmbViewZ = make_float4(unpackViewZ(cb, zr.x), unpackViewZ(cb, zr.y), unpackViewZ(cb, zr.z), unpackViewZ(cb, zr.w));
            float3 vmbVv = reconstructViewPosition(vmbPixelUv, cb.frustumPrev, 1.0f, 0.0f);
            float3 Nv = rotate(cb.worldToViewPrev, N);
            float NoXcurr = dot(N, Xprev - cameraDelta);
            float a = Nv.x * vmbVv.x + Nv.y * vmbVv.y, b = Nv.z * vmbVv.z;
            float4 NoXprev = a * (cb.orthoMode == 0.0f ? vmbViewZ : f4(cb.orthoMode)) + b * vmbViewZ;
            float4 planeDist = fabs4(NoXprev - NoXcurr);

            uint4 idv = gather4(p.prevInternalData, vx, vy);
            float3 d00 = unpackInternalData(idv.x), d10 = unpackInternalData(idv.y), d01 = unpackInternalData(idv.z), d11 = unpackInternalData(idv.w);

            auto occ = [&](float pd, float t, float z, float m) {
                return (pd <= t && inDenoisingRange(cb, z) && compareMaterials(materialID, m, cb.specMinMaterial)) ? 1.0f : 0.0f;
            };
            float4 vmbOcclusion = make_float4(occ(planeDist.x, thr.x, vmbViewZ.x, d00.z), occ(planeDist.y, thr.y, vmbViewZ.y, d10.z), occ(planeDist.z, thr.z, vmbViewZ.z, d01.z),
                                              occ(planeDist.w, thr.w, vmbViewZ.w, d11.z));

            fbits += vmbOcclusion.x * 16.0f;
            fbits += vmbOcclusion.y * 32.0f;
            fbits += vmbOcclusion.z * 64.0f;
            fbits += vmbOcclusion.w * 128.0f;

            vmbOcclusionWeights = bilinearCustomWeights(vmbBilinearFilter, vmbOcclusion);
            vmbSpecAccumSpeed = applyCustomWeights(d00.y, d10.y, d01.y, d11.y, vmbOcclusionWeights);

            float vmbFootprintQuality = sqrt01(applyBilinear(vmbOcclusion.x, vmbOcclusion.y, vmbOcclusion.z, vmbOcclusion.w, vmbBilinearFilter));
            float vmbSpecHistoryConfidence = vmbFootprintQuality;
            if (OPTIONAL && cb.hasHistoryConfidence) vmbSpecHistoryConfidence = fminf(vmbSpecHistoryConfidence, saturate(p.specConfidence.sampleLinear(vmbPixelUv)));
            vmbSpecAccumSpeed *= lerp(vmbSpecHistoryConfidence, 1.0f, 1.0f / (1.0f + vmbSpecAccumSpeed));

            vmbAllowCatRom = sum4(vmbOcclusion) > 3.5f && smbAllowCatRom;
        }

        // How many radians can virtual motion cover?
        float curvatureAngle, lobeHalfAngle;
        {
            float curvatureAngleTan = pixelSize * fabsf(curvature);
            curvatureAngleTan *= fmaxf(vmbPixelsTraveled / fmaxf(NoV, 0.01f), 1.0f);
            curvatureAngleTan *= 2.0f;
            curvatureAngle = atanf(curvatureAngleTan);

            float percentOfVolume = NRD_MAX_PERCENT_OF_LOBE_VOLUME / (1.0f + vmbSpecAccumSpeed);
            float lobeTanHalfAngle = fmaxf(specularLobeTanHalfAngle(roughness, percentOfVolume), NRD_NORMAL_ENCODING_ERROR);
            hitDistanceToLobeSpreadInPixels *= lobeTanHalfAngle;
            lobeHalfAngle = atanf(lobeTanHalfAngle);
        }

        // Virtual motion - confidence: parallax
        float parallaxWeight;
        {
            float hitDistForTrackingPrev = p.prevSpecHitDistForTracking.sampleLinear(vmbPixelUv * resolutionScalePrev);
            float3 XvirtualPrev = getXvirtual(hitDistForTrackingPrev, curvature, X, Xprev, N, V, roughness);
            float2 vmbPixelUvPrev = screenUv(cb.worldToClipPrev, XvirtualPrev);
            vmbPixelUvPrev = materialID == cb.cameraAttachedReflectionMaterialID ? smbPixelUv : vmbPixelUvPrev;

            float r = fminf(hitDistForTracking, hitDistForTrackingPrev) * hitDistanceToLobeSpreadInPixels;
            r *= 0.5f;
            r = fmaxf(r, 0.1f * roughness);
            float d = length((vmbPixelUvPrev - vmbPixelUv) * rectSize);
            parallaxWeight = linearStep(r, 0.0f, d);
        }

        // Virtual motion - confidence: normal
        {
            float normalWeight = encodingAwareNormalWeight(N, xyz(vmbN), lobeHalfAngle, curvatureAngle, 0.0f);
            normalWeight = lerp(1.0f, normalWeight, smoothStep01(vmbPixelsTraveled));
            virtualHistoryConfidence *= normalWeight;
        }

        // Virtual motion - confidence: prev-prev test (one iteration)
        {
            float stepBetweenTaps = fminf(vmbPixelsTraveled * cb.framerateScale, 2.0f) + vmbPixelsTraveled;
            vmbDelta *= rsqrtSafe(dot(vmbDelta, vmbDelta));
            vmbDelta /= rectSizePrev;

            float2 rp = relaxedRoughnessWeightParams(vmbN.w * vmbN.w, cb.roughnessFraction, NRD_ROUGHNESS_SENSITIVITY * 0.3f);
            float2 vmbPixelUvPrev = vmbPixelUv + vmbDelta * stepBetweenTaps;

            int tx, ty;
            stochasticBilinearTexel(rng, vmbPixelUvPrev, rectSizePrev, resolutionScalePrev, p.prevNormalRoughness.w, p.prevNormalRoughness.h, tx, ty);
            float4 prevNR = unpackNormalRoughness(p.prevNormalRoughness.fetchRawClamped(tx, ty));

            float w = encodingAwareNormalWeight(xyz(vmbN), xyz(prevNR), lobeHalfAngle, curvatureAngle * (1.0f + stepBetweenTaps), 0.0f);
            w *= nonExponentialWeight(prevNR.w * prevNR.w, rp.x, rp.y);
            w = lerp(1.0f, w, saturate(stepBetweenTaps));
            w = isInScreenNearest(vmbPixelUvPrev) ? w : 1.0f;
            virtualHistoryConfidence = fminf(virtualHistoryConfidence, w);
        }

        virtualHistoryConfidence *= parallaxWeight;

        // Surface history confidence
        float surfaceHistoryConfidence;
        {
            float a = atanf(smbParallaxInPixelsMax * pixelSize / length(X));
            float nonLinearAccumSpeed = 1.0f / (1.0f + smbSpecAccumSpeed);
            float hPrev;
            if constexpr (FIXED) hPrev = p.historySpec.sampleLinear(smbPixelUv * resolutionScalePrev).w;
            else {   // the same clamped bilinear fetch through the polymorphic accessor
                const float2 uvp = smbPixelUv * resolutionScalePrev;
                const float tx = uvp.x * (float)p.historySpec.w - 0.5f, ty = uvp.y * (float)p.historySpec.h - 0.5f;
                const float fx = floorf(tx), fy = floorf(ty);
                const int x0 = (int)fx, y0 = (int)fy;
                const float a = Sig<MODE>::fetchClamped(p.historySpec, x0, y0).w, b = Sig<MODE>::fetchClamped(p.historySpec, x0 + 1, y0).w;
                const float c = Sig<MODE>::fetchClamped(p.historySpec, x0, y0 + 1).w, d = Sig<MODE>::fetchClamped(p.historySpec, x0 + 1, y0 + 1).w;
                hPrev = lerp(lerp(a, b, tx - fx), lerp(c, d, tx - fx), ty - fy);
            }
            float h = lerp(hPrev, spec.w, nonLinearAccumSpeed) * hitDistNormalization;

            float tana0 = specularLobeTanHalfAngle(roughnessModified, NRD_MAX_PERCENT_OF_LOBE_VOLUME);
            tana0 *= lerp(NoV, 1.0f, roughnessModified);
            tana0 *= nonLinearAccumSpeed;
            tana0 /= hitDistFactor(h, frustumSize) + NRD_EPS;

            float a0 = fmaxf(atanf(tana0), NRD_NORMAL_ENCODING_ERROR);
            float f = linearStep(a0, 0.0f, a);
            surfaceHistoryConfidence = pow01(f, 4.0f);
            f = linearStep(0.8f, 0.9f, roughnessModified);
            surfaceHistoryConfidence = lerp(surfaceHistoryConfidence, 1.0f, f);
        }

        // Limit number of accumulated frames
        float smbSpecAccumSpeed_NoHistoryFix, vmbSpecAccumSpeed_NoHistoryFix;
        {
            float rf = responsiveFactor(cb, roughnessModified);
            float smc = specMagicCurve(roughnessModified);
            float pw = lerp(32.0f, 1.0f, smc) * (1.0f - rf);
            float2 f = lerp(smc, 1.0f, rf) * make_float2(pow01(smbNoN, pw), pow01(vmbNoN, pw));

            float2 maxResponsiveFrameNum = max2(cb.maxAccumulatedFrameNum * f, f2((float)cb.responsiveAccumulationMinAccumulatedFrameNum));
            float2 maxFrameNum = cb.maxAccumulatedFrameNum * make_float2(surfaceHistoryConfidence, virtualHistoryConfidence);
            float2 maxFrameNumNoFix = min2(maxFrameNum, max2(maxResponsiveFrameNum, f2(cb.historyFixFrameNum)));

            smbSpecAccumSpeed_NoHistoryFix = fminf(smbSpecAccumSpeed, maxFrameNumNoFix.x);
            vmbSpecAccumSpeed_NoHistoryFix = fminf(vmbSpecAccumSpeed, maxFrameNumNoFix.y);

            maxFrameNum = min2(maxFrameNum, maxResponsiveFrameNum);
            smbSpecAccumSpeed = fminf(smbSpecAccumSpeed, maxFrameNum.x);
            vmbSpecAccumSpeed = fminf(vmbSpecAccumSpeed, maxFrameNum.y);
        }

        // Virtual history amount
        virtualHistoryAmount = saturate(1.0f + (vmbSpecAccumSpeed - smbSpecAccumSpeed) / (1.0f + 0.5f * fmaxf(vmbSpecAccumSpeed, smbSpecAccumSpeed)));
        if (!smbAllowCatRom || !vmbAllowCatRom) virtualHistoryAmount = step(0.5f, virtualHistoryAmount);

        // Sample history
        float4 specHistory;
        float specFastHistory;
        float4 specShHistory = f4(0.0f);
        {
            float2 uv = lerp(smbPixelUv, vmbPixelUv, virtualHistoryAmount);
            float4 occlusionWeights = lerp(smbOcclusionWeights, vmbOcclusionWeights, virtualHistoryAmount);
            bool allowCatRom = virtualHistoryAmount < 0.5f ? smbAllowCatRom : vmbAllowCatRom;
            HistoryFilter hf(saturate(uv) * rectSizePrev, resourceSizeInvPrev, occlusionWeights, allowCatRom);
            if constexpr (FIXED) {
                specHistory = clampNegativeToZero(hf.color(p.historySpec));
                specFastHistory = fmaxf(hf.bilinear(p.historySpecFast), 0.0f);
            } else {
                specHistory = clampNegativeToZeroM<MODE>(hf.template colorAny<Sig<MODE>>(p.historySpec));
                specFastHistory = fmaxf(hf.template bilinearAny<FastSig<MODE>>(p.historySpecFast), 0.0f);
            }
            if constexpr (SH) specShHistory = hf.bilinear4(p.historySpecSh);
        }

        // Accumulation
        specAccumSpeedCorrected = lerp(smbSpecAccumSpeed_NoHistoryFix, vmbSpecAccumSpeed_NoHistoryFix, virtualHistoryAmount);
        const float specAccumSpeed = lerp(smbSpecAccumSpeed, vmbSpecAccumSpeed, virtualHistoryAmount);
        const float specNonLinearAccumSpeed = checkerboardResolveAccumSpeed(cb, 1.0f / (1.0f + specAccumSpeed), specHasData);

        float4 specResult = mixHistoryAndCurrent(cb, specHistory, spec, specNonLinearAccumSpeed, roughness);
        float4 specShResult = f4(0.0f);
        if constexpr (SH) specShResult = lerp(specShHistory, p.inSpecSh.load(px, py), specNonLinearAccumSpeed);

        // Firefly suppressor
        const float specMaxRelativeIntensity = cb.fireflySuppressorMinRelativeScale + 38.0f / (specAccumSpeed + 1.0f);
        float specAntifireflyFactor = specAccumSpeed * cb.maxBlurRadius * 0.1f;
        specAntifireflyFactor /= 1.0f + specAntifireflyFactor;
        if constexpr (FIXED) {
            float lumaResult = specResult.x;
            float lumaClamped = fminf(lumaResult, specHistory.x * specMaxRelativeIntensity);
            lumaClamped = lerp(lumaResult, lumaClamped, specAntifireflyFactor);
            specResult = changeLuma(specResult, lumaClamped);
            if constexpr (SH) specShResult = rescaleSh(specShResult, lumaClamped);

            float hitDistMaxRelativeIntensity = 1.2f + 1.0f / (specAccumSpeed + 1.0f);
            specResult.w = lerp(specResult.w, fminf(specResult.w, specHistory.w * hitDistMaxRelativeIntensity), specAntifireflyFactor);
        }
        Sig<MODE>::store(p.outSpec, px, py, specResult);
        if constexpr (SH) p.outSpecSh.store(px, py, specShResult);

        {  // Fast history
            float maxFastAccumulatedFrameNum = cb.maxFastAccumulatedFrameNum;
            if (materialID == cb.strandMaterialID) maxFastAccumulatedFrameNum = fmaxf(maxFastAccumulatedFrameNum, cb.maxAccumulatedFrameNum / 5.0f);

            float specHistoryConfidence = lerp(surfaceHistoryConfidence, virtualHistoryConfidence, virtualHistoryAmount);
            float fastNonLinearAccumSpeed = nonLinearAccumSpeedFast(cb, specAccumSpeed, maxFastAccumulatedFrameNum, specHistoryConfidence, specHasData);
            float fastResult = lerp(specFastHistory, lumaOf<MODE>(spec), fastNonLinearAccumSpeed);
            if constexpr (FIXED) {
                float fastClamped = fminf(fastResult, specHistory.x * specMaxRelativeIntensity * 4.0f);
                fastResult = lerp(fastResult, fastClamped, specAntifireflyFactor);
            }
            FastSig<MODE>::store(p.outSpecFast, px, py, fastResult);
        }
    }

    if constexpr (!OCC) storeData2<SIGNAL>(p, px, py, fbits, curvature, virtualHistoryAmount, smbAllowCatRom);   // TA:873-877

    // =============================================================================================== Diffuse
    if constexpr ((SIGNAL & SIGNAL_DIFF) == 0) diffAccumSpeed = 0.0f;  // TA:972-974
    else {
        float diffHistoryConfidence = smbFootprintQuality;
        if (OPTIONAL && cb.hasHistoryConfidence) diffHistoryConfidence = fminf(diffHistoryConfidence, saturate(p.diffConfidence.sampleLinear(smbPixelUv)));
        diffAccumSpeed *= lerp(diffHistoryConfidence, 1.0f, 1.0f / (1.0f + diffAccumSpeed));

        const bool diffHasData = !OPTIONAL || cb.diffCheckerboard == 2u || (((uint32_t)(px ^ py) ^ cb.frameIndex) & 1u) == cb.diffCheckerboard;
        const float4 diff = currentSignal(p.inDiff, cb.diffCheckerboard, diffHasData);

        HistoryFilter hf(saturate(smbPixelUv) * rectSizePrev, resourceSizeInvPrev, smbOcclusionWeights, smbAllowCatRom);
        float4 diffHistory;
        float diffFastHistory;
        if constexpr (FIXED) {
            diffHistory = clampNegativeToZero(hf.color(p.historyDiff));
            diffFastHistory = fmaxf(hf.bilinear(p.historyDiffFast), 0.0f);
        } else {
            diffHistory = clampNegativeToZeroM<MODE>(hf.template colorAny<Sig<MODE>>(p.historyDiff));
            diffFastHistory = fmaxf(hf.template bilinearAny<FastSig<MODE>>(p.historyDiffFast), 0.0f);
        }
        float4 diffShResult = f4(0.0f);
        if constexpr (SH) diffShResult = hf.bilinear4(p.historyDiffSh);

        const float diffNonLinearAccumSpeed = checkerboardResolveAccumSpeed(cb, 1.0f / (1.0f + diffAccumSpeed), diffHasData);
        float4 diffResult = mixHistoryAndCurrent(cb, diffHistory, diff, diffNonLinearAccumSpeed);
        if constexpr (SH) diffShResult = lerp(diffShResult, p.inDiffSh.load(px, py), diffNonLinearAccumSpeed);

        const float diffMaxRelativeIntensity = cb.fireflySuppressorMinRelativeScale + 38.0f / (diffAccumSpeed + 1.0f);
        float diffAntifireflyFactor = diffAccumSpeed * cb.maxBlurRadius * 0.1f;
        diffAntifireflyFactor /= 1.0f + diffAntifireflyFactor;

        if constexpr (FIXED) {
            float lumaResult = diffResult.x;
            float lumaClamped = fminf(lumaResult, diffHistory.x * diffMaxRelativeIntensity);
            lumaClamped = lerp(lumaResult, lumaClamped, diffAntifireflyFactor);
            diffResult = changeLuma(diffResult, lumaClamped);
            if constexpr (SH) diffShResult = rescaleSh(diffShResult, lumaClamped);

            float hitDistMaxRelativeIntensity = 1.2f + 1.0f / (diffAccumSpeed + 1.0f);
            diffResult.w = lerp(diffResult.w, fminf(diffResult.w, diffHistory.w * hitDistMaxRelativeIntensity), diffAntifireflyFactor);
        }
        Sig<MODE>::store(p.outDiff, px, py, diffResult);
        if constexpr (SH) p.outDiffSh.store(px, py, diffShResult);

        float fastNonLinearAccumSpeed = checkerboardResolveAccumSpeed(cb, 1.0f / (1.0f + fminf(diffAccumSpeed, cb.maxFastAccumulatedFrameNum)), diffHasData);
        float fastResult = lerp(diffFastHistory, lumaOf<MODE>(diff), fastNonLinearAccumSpeed);
        if constexpr (FIXED) {
            float fastClamped = fminf(fastResult, diffHistory.x * diffMaxRelativeIntensity * 4.0f);
            fastResult = lerp(fastResult, fastClamped, diffAntifireflyFactor);
        }
        FastSig<MODE>::store(p.outDiffFast, px, py, fastResult);
    }

    storeData1<SIGNAL>(p, px, py, diffAccumSpeed, specAccumSpeedCorrected);
}

void launchReblurTemporalAccumulation(const ReblurConstants& cb, const TemporalAccumulationParams& p, int signal, int mode, Rows rows, cudaStream_t stream) {
    const RowGrid g = rowGrid(rows, cb.rectSizeMinusOne[1] + 1, BLOCK_H);
    if (!g.count) return;
    const dim3 grid((cb.rectSizeMinusOne[0] + BLOCK_W) / BLOCK_W, g.count), block(BLOCK_W, BLOCK_H);
    const bool optional = cb.specCheckerboard != 2u || cb.diffCheckerboard != 2u || cb.hasHistoryConfidence || cb.hasDisocclusionThresholdMix;
    if (mode == MODE_DO) {   // one denoiser, one lobe
        launchK(reblurTemporalAccumulationKernel<true, SIGNAL_DIFF, MODE_DO>, grid, block, 0, stream, cb, p, g.ctaY0);
        return;
    }
    withSignal(signal, [&](auto sig) {
        constexpr int S = decltype(sig)::value;
        if (mode == MODE_OCCLUSION) launchK(reblurTemporalAccumulationKernel<true, S, MODE_OCCLUSION>, grid, block, 0, stream, cb, p, g.ctaY0);
        else if (mode == MODE_SH) {
            if (optional) launchK(reblurTemporalAccumulationKernel<true, S, MODE_SH>, grid, block, 0, stream, cb, p, g.ctaY0);
            else launchK(reblurTemporalAccumulationKernel<false, S, MODE_SH>, grid, block, 0, stream, cb, p, g.ctaY0);
        } else {
            if (optional) launchK(reblurTemporalAccumulationKernel<true, S, MODE_RADIANCE>, grid, block, 0, stream, cb, p, g.ctaY0);
            else launchK(reblurTemporalAccumulationKernel<false, S, MODE_RADIANCE>, grid, block, 0, stream, cb, p, g.ctaY0);
        }
    });
}

}  // namespace nrdk
