// REBLUR HistoryFix and TemporalStabilization on sm_100a (NRD_MODE=RADIANCE; NRD_SIGNAL = DIFF / SPEC / BOTH is a template parameter), plus Clear.
//
// Replaces External/NRD/Shaders/REBLUR_HistoryFix.cs.hlsl:44-506 (sparse 5x5-minus-corners cross-bilateral
// reconstruction for pixels with < historyFixFrameNum frames of history, then 9x9 anti-firefly and 5x5 fast-history
// luma clamps), REBLUR_TemporalStabilization.cs.hlsl:41-318 (3x3 luma moments, antilag, Catmull-Rom fetch of the
// stabilised-luma history, clamp + lerp, final outputs and packed internal data) and Clear.cs.hlsl:18-24.
//
// Mapping to the GPU: the dense luma stencils are the textbook shared-memory case — a CTA stages the clamped
// (32+2B) x (16+2B) tile of both fast-history lumas once (sky texels flagged REBLUR_INVALID), each thread then reads
// its window from shared memory. The sparse reconstruction taps (stride <= 14 px, only on freshly disoccluded pixels)
// and the history fetches stay as L2 gathers.
#include "reblur_common.cuh"

namespace nrdk {

namespace {
enum { DIFF = 0, SPEC = 1 };
constexpr int BLOCK_W = 32, BLOCK_H = 16;
}  // namespace

// ===============================================================================================================
// History fix
// ===============================================================================================================
namespace {
constexpr int HF_BORDER = 4;  // REBLUR_ANTI_FIREFLY_FILTER_RADIUS
constexpr int HF_TILE_W = BLOCK_W + 2 * HF_BORDER, HF_TILE_H = BLOCK_H + 2 * HF_BORDER;

// SH ( NRD_MODE = SH ): the lobe's second RGBA16F is reconstructed with the same weights and rescaled to the clamped luma ( REBLUR_HistoryFix.cs.hlsl:106-108, 135-137, 194-207, 286-294 )
// MODE ( NRD_MODE ): OCCLUSION / DO go through the format-polymorphic Sig / FastSig, their "luma" is the hit distance, there is no anti-firefly clamp
// ( NRD_SUPPORTS_ANTIFIREFLY = 0 ) and OCCLUSION has no hit-distance-for-tracking texture ( REBLUR_HistoryFix.cs.hlsl:311-314 )
template <int LOBE, int SIGNAL, int MODE>
NRD_DEV void historyFixLobe(const ReblurConstants& cb, const HistoryFixParams& p, const float (*sLuma)[HF_TILE_W], const float2 (*sRow)[HF_TILE_H][BLOCK_W], bool tileHasSky,
                            int px, int py, float strideIn, float frameNum,
                            float frameNumAvgNorm, float viewZ, float materialID, float3 N, float roughness, float3 Nv, float3 Xv, float frustumSize, float2 pixelUv) {
    const TexRGBA16F& in = LOBE == DIFF ? p.inDiff : p.inSpec;
    const TexRGBA16F& out = LOBE == DIFF ? p.outDiff : p.outSpec;
    const TexR16F& outFast = LOBE == DIFF ? p.outDiffFast : p.outSpecFast;
    const float MIN_MATERIAL = LOBE == DIFF ? cb.diffMinMaterial : cb.specMinMaterial;
    const float2 rectSize = make_float2(cb.rectSize[0], cb.rectSize[1]);
    const float2 rectSizeInv = make_float2(cb.rectSizeInv[0], cb.rectSizeInv[1]);

    constexpr bool SH = MODE == MODE_SH, FIXED = Sig<MODE>::FIXED;
    float4 v = Sig<MODE>::load(in, px, py);
    float4 sh = f4(0.0f);
    if constexpr (SH) sh = (LOBE == DIFF ? p.inDiffSh : p.inSpecSh).load(px, py);
    const float smc = LOBE == DIFF ? 1.0f : specMagicCurve(roughness);
    const float nonLinearAccumSpeed = 1.0f / (1.0f + frameNum);

    const float hitDistScale = hitDistanceNormalization(viewZ, cb.hitDistSettings, LOBE == DIFF ? 1.0f : roughness);
    float hitDist = v.w * hitDistScale;
    if (LOBE == SPEC && MODE != MODE_OCCLUSION) hitDist = lerp(p.specHitDistForTracking.load(px, py), hitDist, smc);
    const float hdFactor = hitDistFactor(hitDist, frustumSize);
    hitDist = LOBE == DIFF ? v.w : saturate(hitDist / hitDistScale);

    float stride = strideIn;
    stride *= lerp(0.25f + 0.75f * sqrt01(hdFactor), 1.0f, nonLinearAccumSpeed);
    if (LOBE == SPEC) stride *= lerp(0.25f, 1.0f, smc);
    stride = roundNe(stride);

    if (stride != 0.0f) {
        const float normalParam = normalWeightParam(nonLinearAccumSpeed, cb.lobeAngleFraction, LOBE == DIFF ? 1.0f : roughness);
        const float2 geomParams = geometryWeightParams(cb.planeDistSensitivity, frustumSize, Xv, Nv);
        const float2 hitDistParams = hitDistanceWeightParams(hitDist, nonLinearAccumSpeed);
        const float2 roughParams = relaxedRoughnessWeightParams(roughness * roughness, sqrtf(cb.roughnessFraction));

        float sum = 1.0f + frameNum;
        v *= sum;
        sh *= sum;

        for (int j = -2; j <= 2; j++)
            for (int i = -2; i <= 2; i++) {
                if ((i == 0 && j == 0) || (abs(i) + abs(j) == 4)) continue;

                float2 uv = mirrorUv(pixelUv + make_float2((float)i, (float)j) * stride * rectSizeInv);
                float2 posf = uv * rectSize;
                int tx = (int)posf.x, ty = (int)posf.y;

                float zs = unpackViewZ(cb, p.viewZ.load(tx, ty));
                float3 Xvs = reconstructViewPosition(uv, cb.frustum, zs, cb.orthoMode);

                float materialIDs;
                float4 Ns = unpackNormalRoughness(p.normalRoughness.loadRaw(tx, ty), materialIDs);

                float angle = acosApproxPositive(dot(xyz(Ns), N));
                float NoX = dot(Nv, Xvs);

                float w = compareMaterials(materialID, materialIDs, MIN_MATERIAL) ? 1.0f : 0.0f;
                w *= exponentialWeight(angle, normalParam, 0.0f);
                if (LOBE == SPEC) w *= exponentialWeight(Ns.w * Ns.w, roughParams.x, roughParams.y);

                float2 fn = loadData1<SIGNAL>(p, tx, ty);
                w *= 1.0f + (LOBE == DIFF ? fn.x : fn.y);
                w = applyGeometryWeightLast(cb, w, zs, NoX, geomParams);

                float4 smp = Sig<MODE>::load(in, tx, ty);
                smp = w == 0.0f ? f4(0.0f) : smp;
                w *= exponentialWeight(smp.w, hitDistParams.x, hitDistParams.y);

                sum += w;
                v += smp * w;
                if constexpr (SH) {
                    float4 shs = (LOBE == DIFF ? p.inDiffSh : p.inSpecSh).load(tx, ty);
                    shs = w == 0.0f ? f4(0.0f) : shs;
                    sh += shs * w;
                }
            }
        v *= positiveRcp(sum);
        sh *= positiveRcp(sum);
    }

    float luma = lumaOf<MODE>(v);

    float f = frameNumAvgNorm;
    if (LOBE == SPEC) f = lerp(1.0f, f, smc);

    const int sx = threadIdx.x + HF_BORDER, sy = threadIdx.y + HF_BORDER;
    float fastCenter = lerp(luma, sLuma[sy][sx], f);
    FastSig<MODE>::store(outFast, px, py, fastCenter);

    // Local variance: 5x5 for the fast-history clamp, 9x9 minus the central 3x3 for the anti-firefly clamp
    float fastM1 = fastCenter, fastM2 = fastCenter * fastCenter;
    float antiFireflyM1 = 0.0f, antiFireflyM2 = 0.0f;
    if (!tileHasSky) {
        // Separable box sums: the CTA already reduced every tile row over the 3 / 5 / 9 wide windows (packed { v, v^2 });
        // each thread adds 3 + 5 + 9 row sums with FADD2 instead of walking 80 texels. The centre texel enters the
        // reference's sums as `fastCenter`, not as its stored value, hence the "- centre" terms.
        const int x = threadIdx.x, y = threadIdx.y + HF_BORDER;
        float2 s3 = sRow[0][y][x], s5 = sRow[1][y][x], s9 = sRow[2][y][x];
#pragma unroll
        for (int j = 1; j <= HF_BORDER; j++) {
            if (j <= 1) s3 = __fadd2_rn(s3, __fadd2_rn(sRow[0][y - j][x], sRow[0][y + j][x]));
            if (j <= 2) s5 = __fadd2_rn(s5, __fadd2_rn(sRow[1][y - j][x], sRow[1][y + j][x]));
            s9 = __fadd2_rn(s9, __fadd2_rn(sRow[2][y - j][x], sRow[2][y + j][x]));
        }
        const float vc = sLuma[sy][sx];
        fastM1 += s5.x - vc;
        fastM2 += s5.y - vc * vc;
        antiFireflyM1 = s9.x - s3.x;
        antiFireflyM2 = s9.y - s3.y;
    } else {
        // Tiles touching the sky keep the reference's per-texel walk: invalid texels are replaced by this pixel's own value
#pragma unroll 1
        for (int j = -HF_BORDER; j <= HF_BORDER; j++)
#pragma unroll
            for (int i = -HF_BORDER; i <= HF_BORDER; i++) {
                if (i == 0 && j == 0) continue;
                float d = sLuma[sy + j][sx + i];
                d = d == REBLUR_INVALID ? fastCenter : d;
                if (abs(i) <= 2 && abs(j) <= 2) {
                    fastM1 += d;
                    fastM2 += d * d;
                }
                if (!(abs(i) <= 1 && abs(j) <= 1)) {
                    antiFireflyM1 += d;
                    antiFireflyM2 += d * d;
                }
            }
    }

    if (FIXED && cb.antiFirefly != 0.0f) {
        const float invNorm = 1.0f / ((HF_BORDER * 2 + 1) * (HF_BORDER * 2 + 1) - 3 * 3);
        antiFireflyM1 *= invNorm;
        antiFireflyM2 *= invNorm;
        float sigma = stdDev(antiFireflyM1, antiFireflyM2) * 2.0f;
        luma = clampf(luma, antiFireflyM1 - sigma, antiFireflyM1 + sigma);
    }
    {
        fastM1 *= 1.0f / 25.0f;
        fastM2 *= 1.0f / 25.0f;
        float scale = cb.fastHistoryClampingSigmaScale;
        if (LOBE == SPEC && materialID == cb.strandMaterialID) scale = fmaxf(scale, 3.0f);
        float sigma = stdDev(fastM1, fastM2) * scale;
        float lumaClamped = clampf(luma, fastM1 - sigma, fastM1 + sigma);
        luma = lerp(lumaClamped, luma, 1.0f / (1.0f + (cb.maxFastAccumulatedFrameNum < cb.maxAccumulatedFrameNum ? 1.0f : 0.0f) * frameNum * 2.0f));
    }

    Sig<MODE>::store(out, px, py, changeLumaM<MODE>(v, luma));
    if constexpr (SH) (LOBE == DIFF ? p.outDiffSh : p.outSpecSh).store(px, py, rescaleSh(sh, luma));
}
}  // namespace

#ifndef HF_MIN_BLOCKS
#    define HF_MIN_BLOCKS 3  // 512-thread CTAs: 40 regs (3 CTAs / SM) 153 us vs 156 us at 57 regs (2 CTAs)
#endif
template <int SIGNAL, int MODE>
__global__ void __launch_bounds__(BLOCK_W* BLOCK_H, HF_MIN_BLOCKS) reblurHistoryFixKernel(const __grid_constant__ ReblurConstants cb, const __grid_constant__ HistoryFixParams p, int quads, int ctaY0) {
    pdlEntry();
    constexpr bool HAS_DIFF = (SIGNAL & SIGNAL_DIFF) != 0, HAS_SPEC = (SIGNAL & SIGNAL_SPEC) != 0;
    __shared__ float sDiffLuma[HAS_DIFF ? HF_TILE_H : 1][HF_TILE_W];
    __shared__ float sSpecLuma[HAS_SPEC ? HF_TILE_H : 1][HF_TILE_W];
    // row sums of { v, v^2 } over windows of 3 / 5 / 9 texels centred on the 32 interior columns, per lobe
    __shared__ float2 sDiffRow[3][HAS_DIFF ? HF_TILE_H : 1][BLOCK_W];
    __shared__ float2 sSpecRow[3][HAS_SPEC ? HF_TILE_H : 1][BLOCK_W];

    const int2 cta = ctaTile<2>(ctaY0);
    const int px = cta.x * BLOCK_W + threadIdx.x, py = cta.y * BLOCK_H + threadIdx.y;
    const int tid = threadIdx.y * BLOCK_W + threadIdx.x;
    int sawSky = 0;
    {
        const int baseX = cta.x * BLOCK_W - HF_BORDER, baseY = cta.y * BLOCK_H - HF_BORDER;
        for (int i = tid; i < HF_TILE_W * HF_TILE_H; i += BLOCK_W * BLOCK_H) {
            int sx = i % HF_TILE_W, sy = i / HF_TILE_W;
            int gx = clampi(baseX + sx, 0, cb.rectSizeMinusOne[0]), gy = clampi(baseY + sy, 0, cb.rectSizeMinusOne[1]);
            // the coordinates are clamped to the rect: plain fetches, issued together with the viewZ load instead of behind its sky test
            float dFast = 0.0f, sFast = 0.0f;
            if constexpr (HAS_DIFF) dFast = FastSig<MODE>::fetch(p.inDiffFast, gx, gy);
            if constexpr (HAS_SPEC) sFast = FastSig<MODE>::fetch(p.inSpecFast, gx, gy);
            bool sky = !inDenoisingRange(cb, unpackViewZ(cb, p.viewZ.fetch(gx, gy)));
            sawSky |= sky ? 1 : 0;
            if constexpr (HAS_DIFF) sDiffLuma[sy][sx] = sky ? REBLUR_INVALID : dFast;
            if constexpr (HAS_SPEC) sSpecLuma[sy][sx] = sky ? REBLUR_INVALID : sFast;
        }
    }
    const bool tileHasSky = __syncthreads_or(sawSky) != 0;  // also the barrier that publishes the tile
    if (!tileHasSky) {
        for (int i = tid; i < HF_TILE_H * BLOCK_W; i += BLOCK_W * BLOCK_H) {
            const int x = i % BLOCK_W, y = i / BLOCK_W, c = x + HF_BORDER;
            auto sq = [](float v) { return make_float2(v, v * v); };
            auto rowSums = [&](const float(*t)[HF_TILE_W], auto* r) {
                float2 a = __fadd2_rn(sq(t[y][c]), __fadd2_rn(sq(t[y][c - 1]), sq(t[y][c + 1])));
                r[0][y][x] = a;
                a = __fadd2_rn(a, __fadd2_rn(sq(t[y][c - 2]), sq(t[y][c + 2])));
                r[1][y][x] = a;
                a = __fadd2_rn(a, __fadd2_rn(sq(t[y][c - 3]), sq(t[y][c + 3])));
                a = __fadd2_rn(a, __fadd2_rn(sq(t[y][c - 4]), sq(t[y][c + 4])));
                r[2][y][x] = a;
            };
            if constexpr (HAS_DIFF) rowSums(sDiffLuma, sDiffRow);
            if constexpr (HAS_SPEC) rowSums(sSpecLuma, sSpecRow);
        }
        __syncthreads();
    }

    // Quad exchange first (HistoryFix.cs.hlsl:56-74): all lanes stay until it is done
    const bool skyTile = p.tiles.load(px >> 4, py >> 4) != 0.0f;
    float2 frameNum = loadData1<SIGNAL>(p, px, py);
    const float viewZ = unpackViewZ(cb, p.viewZ.load(px, py));
    if (!inDenoisingRange(cb, viewZ)) frameNum = f2(REBLUR_MAX_ACCUM_FRAME_NUM);
    float2 stride = make_float2(frameNum.x < cb.historyFixFrameNum ? 1.0f : 0.0f, frameNum.y < cb.historyFixFrameNum ? 1.0f : 0.0f);
    if (quads) {
        float2 d10 = make_float2(__shfl_xor_sync(0xFFFFFFFFu, stride.x, 1), __shfl_xor_sync(0xFFFFFFFFu, stride.y, 1));
        float2 d01 = make_float2(__shfl_xor_sync(0xFFFFFFFFu, stride.x, 2), __shfl_xor_sync(0xFFFFFFFFu, stride.y, 2));
        stride = min2(stride, (d10 + d01 + stride) / 3.0f);
    }
    if (skyTile || !inDenoisingRange(cb, viewZ) || px > cb.rectSizeMinusOne[0] || py > cb.rectSizeMinusOne[1]) return;

    float materialID;
    const float4 nr = unpackNormalRoughness(p.normalRoughness.loadRaw(px, py), materialID);
    const float3 N = xyz(nr);
    const float roughness = nr.w;

    const float frustumSize = frustumSizeAt(cb.minRectDimMulUnproject, cb.orthoMode, viewZ);
    const float2 pixelUv = make_float2(px + 0.5f, py + 0.5f) * make_float2(cb.rectSizeInv[0], cb.rectSizeInv[1]);
    const float3 Xv = reconstructViewPosition(pixelUv, cb.frustum, viewZ, cb.orthoMode);
    const float3 Nv = rotateInverse(cb.viewToWorld, N);

    const float invHistoryFixFrameNum = 1.0f / fmaxf(cb.historyFixFrameNum, NRD_EPS);
    const float2 frameNumAvgNorm = saturate(frameNum * invHistoryFixFrameNum);

    stride /= 1.0f + 1.0f;
    stride *= 2.0f / 2.0f;
    stride *= materialID == cb.historyFixAlternatePixelStrideMaterialID ? cb.historyFixAlternatePixelStride : cb.historyFixBasePixelStride;

    if constexpr (HAS_DIFF) historyFixLobe<DIFF, SIGNAL, MODE>(cb, p, sDiffLuma, sDiffRow, tileHasSky, px, py, stride.x, frameNum.x, frameNumAvgNorm.x, viewZ, materialID, N, roughness, Nv, Xv, frustumSize, pixelUv);
    if constexpr (HAS_SPEC) historyFixLobe<SPEC, SIGNAL, MODE>(cb, p, sSpecLuma, sSpecRow, tileHasSky, px, py, stride.y, frameNum.y, frameNumAvgNorm.y, viewZ, materialID, N, roughness, Nv, Xv, frustumSize, pixelUv);
}

// ===============================================================================================================
// Temporal stabilization
// ===============================================================================================================
namespace {
constexpr int TS_BORDER = 1;
constexpr int TS_TILE_W = BLOCK_W + 2 * TS_BORDER, TS_TILE_H = BLOCK_H + 2 * TS_BORDER;

NRD_DEV void lumaMoments3x3(const float (*sLuma)[TS_TILE_W], float& luma, float& m1, float& sigma) {
    const int sx = threadIdx.x + TS_BORDER, sy = threadIdx.y + TS_BORDER;
    luma = sLuma[sy][sx];
    m1 = luma;
    float m2 = luma * luma;
#pragma unroll
    for (int j = 0; j <= 2; j++)
#pragma unroll
        for (int i = 0; i <= 2; i++) {
            if (i == 1 && j == 1) continue;
            float d = sLuma[threadIdx.y + j][threadIdx.x + i];
            d = d == REBLUR_INVALID ? luma : d;
            m1 += d;
            m2 += d * d;
        }
    m1 /= 9.0f;
    m2 /= 9.0f;
    sigma = stdDev(m1, m2);
}
}  // namespace

#ifndef TS_MIN_BLOCKS
#    define TS_MIN_BLOCKS 3  // 512-thread CTAs: 40 regs + a small spill (3 CTAs / SM) beats 64 regs (2 CTAs) by 12 % on B200
#endif
// MODE: RADIANCE / SH / DO ( the occlusion denoisers have no stabilization pass ); DO stabilizes the hit distance in .w as its "luma"
template <int SIGNAL, int MODE>
__global__ void __launch_bounds__(BLOCK_W* BLOCK_H, TS_MIN_BLOCKS) reblurTemporalStabilizationKernel(const __grid_constant__ ReblurConstants cb,
                                                                                       const __grid_constant__ TemporalStabilizationParams p, int ctaY0) {
    pdlEntry();
    constexpr bool HAS_DIFF = (SIGNAL & SIGNAL_DIFF) != 0, HAS_SPEC = (SIGNAL & SIGNAL_SPEC) != 0;
    __shared__ float sDiffLuma[HAS_DIFF ? TS_TILE_H : 1][TS_TILE_W];
    __shared__ float sSpecLuma[HAS_SPEC ? TS_TILE_H : 1][TS_TILE_W];
    constexpr bool SH = MODE == MODE_SH, FIXED = Sig<MODE>::FIXED;

    const int2 cta = ctaTile<5>(ctaY0);
    const int px = cta.x * BLOCK_W + threadIdx.x, py = cta.y * BLOCK_H + threadIdx.y;
    {
        const int baseX = cta.x * BLOCK_W - TS_BORDER, baseY = cta.y * BLOCK_H - TS_BORDER;
        const int tid = threadIdx.y * BLOCK_W + threadIdx.x;
        for (int i = tid; i < TS_TILE_W * TS_TILE_H; i += BLOCK_W * BLOCK_H) {
            int sx = i % TS_TILE_W, sy = i / TS_TILE_W;
            int gx = clampi(baseX + sx, 0, cb.rectSizeMinusOne[0]), gy = clampi(baseY + sy, 0, cb.rectSizeMinusOne[1]);
            // clamped coordinates: plain fetches of the luma half-words, issued together with the viewZ load
            float dLuma = 0.0f, sLuma = 0.0f;
            if constexpr (FIXED) {
                if constexpr (HAS_DIFF) dLuma = __half2float(__ushort_as_half(__ldg(p.inDiff.template ptr<unsigned short>(gx * 4, gy * 4) )));
                if constexpr (HAS_SPEC) sLuma = __half2float(__ushort_as_half(__ldg(p.inSpec.template ptr<unsigned short>(gx * 4, gy * 4) )));
            } else {
                if constexpr (HAS_DIFF) dLuma = lumaOf<MODE>(Sig<MODE>::fetch(p.inDiff, gx, gy));
                if constexpr (HAS_SPEC) sLuma = lumaOf<MODE>(Sig<MODE>::fetch(p.inSpec, gx, gy));
            }
            bool sky = !inDenoisingRange(cb, unpackViewZ(cb, p.viewZ.fetch(gx, gy)));
            if constexpr (HAS_DIFF) sDiffLuma[sy][sx] = sky ? REBLUR_INVALID : dLuma;
            if constexpr (HAS_SPEC) sSpecLuma[sy][sx] = sky ? REBLUR_INVALID : sLuma;
        }
    }
    __syncthreads();

    if (p.tiles.load(px >> 4, py >> 4) != 0.0f || px > cb.rectSizeMinusOne[0] || py > cb.rectSizeMinusOne[1]) return;
    const float viewZ = unpackViewZ(cb, p.viewZ.load(px, py));
    if (!inDenoisingRange(cb, viewZ)) return;

    const float2 rectSizeInv = make_float2(cb.rectSizeInv[0], cb.rectSizeInv[1]);
    const float2 rectSizePrev = make_float2(cb.rectSizePrev[0], cb.rectSizePrev[1]);
    const float2 resourceSizeInvPrev = make_float2(cb.resourceSizeInvPrev[0], cb.resourceSizeInvPrev[1]);
    const float3 cameraDelta = make_float3(cb.cameraDelta[0], cb.cameraDelta[1], cb.cameraDelta[2]);

    const float2 pixelUv = make_float2(px + 0.5f, py + 0.5f) * rectSizeInv;
    const float3 Xv = reconstructViewPosition(pixelUv, cb.frustum, viewZ, cb.orthoMode);
    const float3 X = rotate(cb.viewToWorld, Xv);

    float4 mvRaw = p.mv.load(px, py);
    float3 mv = make_float3(mvRaw.x * cb.mvScale[0], mvRaw.y * cb.mvScale[1], mvRaw.z * cb.mvScale[2]);
    float3 Xprev = X;
    float2 smbPixelUv = pixelUv + xy(mv);
    if (cb.mvScale[3] == 0.0f) {
        if (cb.mvScale[2] == 0.0f) mv.z = affine(cb.worldToViewPrev, X).z - viewZ;
        float3 Xvprevlocal = reconstructViewPosition(smbPixelUv, cb.frustumPrev, viewZ + mv.z, cb.orthoMode);
        Xprev = rotateInverse(cb.worldToViewPrev, Xvprevlocal) + cameraDelta;
    } else {
        Xprev += mv;
        smbPixelUv = screenUv(cb.worldToClipPrev, Xprev);
    }

    float materialID;
    const float4 nr = unpackNormalRoughness(p.normalRoughness.loadRaw(px, py), materialID);
    const float3 N = xyz(nr);
    const float roughness = nr.w;

    uint32_t bits;
    bool smbAllowCatRom;
    float2 data1 = loadData1<SIGNAL>(p, px, py);
    const float2 data2 = loadData2<SIGNAL>(p, px, py, bits, smbAllowCatRom);

    const Bilinear smbBilinearFilter = getBilinearFilter(smbPixelUv, rectSizePrev);
    const float4 smbOcclusion = make_float4((bits & 1u) != 0, (bits & 2u) != 0, (bits & 4u) != 0, (bits & 8u) != 0);
    const float4 smbOcclusionWeights = bilinearCustomWeights(smbBilinearFilter, smbOcclusion);
    const float smbFootprintQuality = sqrt01(applyBilinear(smbOcclusion.x, smbOcclusion.y, smbOcclusion.z, smbOcclusion.w, smbBilinearFilter));

    // ---- Diffuse ----
    if constexpr (HAS_DIFF) {
        float luma, m1, sigma;
        lumaMoments3x3(sDiffLuma, luma, m1, sigma);
        if (data1.x < cb.historyFixFrameNum) luma = fminf(luma, m1 * (1.2f + 1.0f / (1.0f + data1.x)));

        HistoryFilter hf(saturate(smbPixelUv) * rectSizePrev, resourceSizeInvPrev, smbOcclusionWeights, smbAllowCatRom);
        float lumaHistory = fmaxf(hf.color(p.historyDiffLuma), 0.0f);

        float antilag = computeAntilag(cb, lumaHistory, m1, sigma, smbFootprintQuality * data1.x);
        float minAccumSpeed = fminf(data1.x, cb.historyFixFrameNum);
        data1.x = lerp(minAccumSpeed, data1.x, antilag);

        float2 params = temporalAccumulationParams(cb, smbFootprintQuality, data1.x, antilag);
        float historyWeight = params.x;
        historyWeight *= pixelUv.x >= cb.splitScreen ? 1.0f : 0.0f;
        historyWeight *= smbPixelUv.x >= cb.splitScreenPrev ? 1.0f : 0.0f;

        float s = sigma * params.y;
        lumaHistory = clampf(lumaHistory, m1 - s, m1 + s);
        float lumaStabilized = lerp(luma, lumaHistory, fminf(historyWeight, cb.stabilizationStrength));

        float4 diff = changeLumaM<MODE>(Sig<MODE>::load(p.inDiff, px, py), lumaStabilized);
        diff.w = cb.returnHistoryLengthInsteadOfOcclusion ? data1.x : diff.w;
        Sig<MODE>::store(p.outDiff, px, py, diff);
        p.outDiffLuma.store(px, py, lumaStabilized);
        if constexpr (SH) p.outDiffSh.store(px, py, rescaleSh(p.inDiffSh.load(px, py), lumaStabilized));  // REBLUR_TemporalStabilization.cs.hlsl:175-187
    }

    // ---- Specular ----
    if constexpr (HAS_SPEC) {
        float luma, m1, sigma;
        lumaMoments3x3(sSpecLuma, luma, m1, sigma);
        if (data1.y < cb.historyFixFrameNum) luma = fminf(luma, m1 * (1.2f + 1.0f / (1.0f + data1.y)));

        const float hitDistForTracking = p.specHitDistForTracking.load(px, py);
        const float virtualHistoryAmount = data2.x, curvature = data2.y;

        const float3 V = viewVector(cb, X);
        const float3 Xvirtual = getXvirtual(hitDistForTracking, curvature, X, Xprev, N, V, roughness);
        float2 vmbPixelUv = screenUv(cb.worldToClipPrev, Xvirtual);
        vmbPixelUv = materialID == cb.cameraAttachedReflectionMaterialID ? pixelUv : vmbPixelUv;

        const Bilinear vmbBilinearFilter = getBilinearFilter(vmbPixelUv, rectSizePrev);
        const float4 vmbOcclusion = make_float4((bits & 16u) != 0, (bits & 32u) != 0, (bits & 64u) != 0, (bits & 128u) != 0);
        const float4 vmbOcclusionWeights = bilinearCustomWeights(vmbBilinearFilter, vmbOcclusion);
        const bool vmbAllowCatRom = sum4(vmbOcclusion) > 3.5f && smbAllowCatRom;
        const float vmbFootprintQuality = sqrt01(applyBilinear(vmbOcclusion.x, vmbOcclusion.y, vmbOcclusion.z, vmbOcclusion.w, vmbBilinearFilter));

        float2 uv = lerp(smbPixelUv, vmbPixelUv, virtualHistoryAmount);
        float4 occlusionWeights = lerp(smbOcclusionWeights, vmbOcclusionWeights, virtualHistoryAmount);
        bool allowCatRom = virtualHistoryAmount < 0.5f ? smbAllowCatRom : vmbAllowCatRom;

        HistoryFilter hf(saturate(uv) * rectSizePrev, resourceSizeInvPrev, occlusionWeights, allowCatRom);
        float lumaHistory = fmaxf(hf.color(p.historySpecLuma), 0.0f);

        float footprintQuality = lerp(smbFootprintQuality, vmbFootprintQuality, virtualHistoryAmount);
        float antilag = computeAntilag(cb, lumaHistory, m1, sigma, footprintQuality * data1.y);
        float minAccumSpeed = fminf(data1.y, cb.historyFixFrameNum);
        data1.y = lerp(minAccumSpeed, data1.y, antilag);

        float2 params = temporalAccumulationParams(cb, footprintQuality, data1.y, antilag);
        float historyWeight = params.x;
        historyWeight *= pixelUv.x >= cb.splitScreen ? 1.0f : 0.0f;
        historyWeight *= virtualHistoryAmount != 1.0f ? (smbPixelUv.x >= cb.splitScreenPrev ? 1.0f : 0.0f) : 1.0f;
        historyWeight *= virtualHistoryAmount != 0.0f ? (vmbPixelUv.x >= cb.splitScreenPrev ? 1.0f : 0.0f) : 1.0f;

        float rf = responsiveFactor(cb, roughness);
        float smc = specMagicCurve(roughness);
        float acceleration = lerp(smc, 1.0f, 0.5f + rf * 0.5f);
        if (materialID == cb.strandMaterialID) acceleration = fminf(acceleration, 0.5f);
        historyWeight *= acceleration;

        float s = sigma * params.y;
        lumaHistory = clampf(lumaHistory, m1 - s, m1 + s);
        float lumaStabilized = lerp(luma, lumaHistory, fminf(historyWeight, cb.stabilizationStrength));

        float4 spec = changeLumaM<MODE>(Sig<MODE>::load(p.inSpec, px, py), lumaStabilized);
        spec.w = cb.returnHistoryLengthInsteadOfOcclusion ? data1.y : spec.w;
        Sig<MODE>::store(p.outSpec, px, py, spec);
        p.outSpecLuma.store(px, py, lumaStabilized);
        if constexpr (SH) p.outSpecSh.store(px, py, rescaleSh(p.inSpecSh.load(px, py), lumaStabilized));  // REBLUR_TemporalStabilization.cs.hlsl:302-314
    }

    p.outInternalData.store(px, py, packInternalData(cb, data1.x, data1.y, materialID));
}

// ===============================================================================================================
// Clear: zero a whole texture (any format) with 16-byte stores where the row allows it
// ===============================================================================================================
__global__ void clearKernel(uint8_t* data, int rowBytes, int height, int pitch) {
    pdlEntry();
    const int y = blockIdx.y;
    uint8_t* row = data + (size_t)y * pitch;
    const int vecs = (((uintptr_t)row & 15) == 0) ? rowBytes / 16 : 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < vecs; i += gridDim.x * blockDim.x) reinterpret_cast<uint4*>(row)[i] = make_uint4(0, 0, 0, 0);
    for (int i = vecs * 16 + blockIdx.x * blockDim.x + threadIdx.x; i < rowBytes; i += gridDim.x * blockDim.x) row[i] = 0;
}

void launchClear(void* data, int rowBytes, int height, int pitch, cudaStream_t stream) {
    if (rowBytes <= 0 || height <= 0) return;
    int blocksX = (rowBytes / 16 + 255) / 256;
    if (blocksX < 1) blocksX = 1;
    if (blocksX > 8) blocksX = 8;
    launchK(clearKernel, dim3(blocksX, height), 256, 0, stream, (uint8_t*)data, rowBytes, height, pitch);
}

void launchReblurHistoryFix(const ReblurConstants& cb, const HistoryFixParams& p, int signal, int mode, bool quads, Rows rows, cudaStream_t stream) {
    const RowGrid g = rowGrid(rows, cb.rectSizeMinusOne[1] + 1, BLOCK_H);
    if (!g.count) return;
    const dim3 grid((cb.rectSizeMinusOne[0] + BLOCK_W) / BLOCK_W, g.count), block(BLOCK_W, BLOCK_H);
    if (mode == MODE_DO) {
        launchK(reblurHistoryFixKernel<SIGNAL_DIFF, MODE_DO>, grid, block, 0, stream, cb, p, quads ? 1 : 0, g.ctaY0);
        return;
    }
    withSignal(signal, [&](auto sig) {
        constexpr int S = decltype(sig)::value;
        if (mode == MODE_SH) launchK(reblurHistoryFixKernel<S, MODE_SH>, grid, block, 0, stream, cb, p, quads ? 1 : 0, g.ctaY0);
        else if (mode == MODE_OCCLUSION) launchK(reblurHistoryFixKernel<S, MODE_OCCLUSION>, grid, block, 0, stream, cb, p, quads ? 1 : 0, g.ctaY0);
        else launchK(reblurHistoryFixKernel<S, MODE_RADIANCE>, grid, block, 0, stream, cb, p, quads ? 1 : 0, g.ctaY0);
    });
}
void launchReblurTemporalStabilization(const ReblurConstants& cb, const TemporalStabilizationParams& p, int signal, int mode, Rows rows, cudaStream_t stream) {
    const RowGrid g = rowGrid(rows, cb.rectSizeMinusOne[1] + 1, BLOCK_H);
    if (!g.count) return;
    const dim3 grid((cb.rectSizeMinusOne[0] + BLOCK_W) / BLOCK_W, g.count), block(BLOCK_W, BLOCK_H);
    if (mode == MODE_DO) {
        launchK(reblurTemporalStabilizationKernel<SIGNAL_DIFF, MODE_DO>, grid, block, 0, stream, cb, p, g.ctaY0);
        return;
    }
    withSignal(signal, [&](auto sig) {
        constexpr int S = decltype(sig)::value;
        if (mode == MODE_SH) launchK(reblurTemporalStabilizationKernel<S, MODE_SH>, grid, block, 0, stream, cb, p, g.ctaY0);
        else launchK(reblurTemporalStabilizationKernel<S, MODE_RADIANCE>, grid, block, 0, stream, cb, p, g.ctaY0);
    });
}

}  // namespace nrdk
