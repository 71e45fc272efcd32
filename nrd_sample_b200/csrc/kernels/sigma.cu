// SIGMA_SHADOW / SIGMA_SHADOW_TRANSLUCENCY passes on sm_100a (kernels templated on TRANSLUCENCY: SIGMA_TYPE = float / float4): ClassifyTiles, SmoothTiles, Copy, Blur (first pass / post-blur),
// TemporalStabilization, SplitScreen. One kernel per reference dispatch, one thread per pixel, CTA = 32x8 pixels for the
// per-pixel passes (a 32-pixel row per warp: coalesced R16F / R32F / R8 rows), 5x5 neighbourhoods through shared memory.
//
// Reference (External/NRD/Shaders): SIGMA_ClassifyTiles.cs.hlsl:24-91, SIGMA_SmoothTiles.cs.hlsl:21-58,
// SIGMA_Copy.cs.hlsl:19-32, SIGMA_Blur.cs.hlsl:21-286, SIGMA_TemporalStabilization.cs.hlsl:21-236,
// SIGMA_SplitScreen.cs.hlsl:21-45. The reference groups are 8x16 (16x16 for the tile passes); the CUDA grid is laid out
// differently but every pixel computes the same function of the same texels.
#include <string>
#include <type_traits>

#include "../../../include/nrd_b200.h"
#include "../../../include/nrdcu.h"
#include "../pipeline_key.h"
#include "sigma_common.cuh"
#include "pairmath.cuh"

namespace nrdk {

namespace {

constexpr int BLOCK_W = 32, BLOCK_H = 8;
constexpr int TILE_W = BLOCK_W + 2 * SIGMA_BORDER, TILE_H = BLOCK_H + 2 * SIGMA_BORDER;

// g_Special8 (Common.hlsli:207-218): { offset.xy, distance } — compile-time constants, so the unrolled tap loop folds the zero components and the Gaussian weights
__device__ constexpr float kSpecial8[8][3] = {{-1.0f, 0.0f, 1.0f},
                                              {0.0f, 1.0f, 1.0f},
                                              {1.0f, 0.0f, 1.0f},
                                              {0.0f, -1.0f, 1.0f},
                                              {-0.25f * 1.41421356237309504880f, 0.25f * 1.41421356237309504880f, 0.5f},
                                              {0.25f * 1.41421356237309504880f, 0.25f * 1.41421356237309504880f, 0.5f},
                                              {0.25f * 1.41421356237309504880f, -0.25f * 1.41421356237309504880f, 0.5f},
                                              {-0.25f * 1.41421356237309504880f, -0.25f * 1.41421356237309504880f, 0.5f}};

// GetGaussianWeight( r ) = exp( -0.66 r^2 ) ( Common.hlsli:346-349 ) at the radii the 5x5 kernel and the 8 sparse taps use, indexed by 4 r^2 = dx^2 + dy^2 of the
// 5x5 offsets ( r = length( offset / 2 ) ): the fp32 values expf gives, so that no tap evaluates an exponential
__device__ constexpr float kGaussian5x5[9] = {1.0f, 0.84789371f, 0.71892375f, 0.0f, 0.51685131f, 0.43823496f, 0.0f, 0.0f, 0.26713529f};

NRD_DEV float applyGeometryWeightLast(const SigmaConstants& cb, float w, float z, float NoX, float2 params) {
    w *= nonExponentialWeight(NoX, params.x, params.y);
    return !sigmaInRange(cb, z) ? 0.0f : w;
}

// ---------------------------------------------------------------------------------------------------------------
// One WARP per 16x16 tile ( 8 tiles per CTA ): every lane takes 8 consecutive pixels of one row — two 128-bit loads of viewZ, one of penumbra — and the three
// 9-bit counters and the InterlockedMax of the reference's s_Mask / s_Radius become warp reductions: no shared memory, no barrier. VEC = the rows of both
// textures are 16-byte aligned ( decided on the host ); otherwise, and for the tiles hanging over the edge of the frame, the pixels are read one by one.
template <bool TR, bool VEC>
__global__ void __launch_bounds__(256) sigmaClassifyTilesKernel(const __grid_constant__ SigmaConstants cb, const __grid_constant__ SigmaClassifyTilesParams p, int tilesW, int tilesH) {
    pdlEntry();
    const int lane = threadIdx.x & 31;
    const int tx = blockIdx.x * 8 + (threadIdx.x >> 5), ty = blockIdx.y;
    if (tx >= tilesW) return;
    const int px0 = tx * 16 + (lane & 1) * 8, py = ty * 16 + (lane >> 1);
    float h[8], z[8];
    float4 st[8];
    const bool whole = px0 + 8 <= p.penumbra.w && py < p.penumbra.h && px0 + 8 <= p.viewZ.w && py < p.viewZ.h && (!TR || (px0 + 8 <= p.translucency.w && py < p.translucency.h));
    if (VEC && whole) {
        const uint4 hv = __ldg(reinterpret_cast<const uint4*>(p.penumbra.ptr<unsigned short>(px0, py)));
        const float4 z0 = __ldg(reinterpret_cast<const float4*>(p.viewZ.ptr<float>(px0, py))), z1 = __ldg(reinterpret_cast<const float4*>(p.viewZ.ptr<float>(px0 + 4, py)));
        const uint32_t hw[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&hw[i]));
            h[2 * i] = f.x;
            h[2 * i + 1] = f.y;
        }
        z[0] = z0.x; z[1] = z0.y; z[2] = z0.z; z[3] = z0.w; z[4] = z1.x; z[5] = z1.y; z[6] = z1.z; z[7] = z1.w;
        if (TR) {
            const uint4 t0 = __ldg(reinterpret_cast<const uint4*>(p.translucency.ptr<uint32_t>(px0, py))), t1 = __ldg(reinterpret_cast<const uint4*>(p.translucency.ptr<uint32_t>(px0 + 4, py)));
            const uint32_t tw[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
#pragma unroll
            for (int i = 0; i < 8; i++) st[i] = make_float4((float)(tw[i] & 255u) / 255.0f, (float)((tw[i] >> 8) & 255u) / 255.0f, (float)((tw[i] >> 16) & 255u) / 255.0f, (float)(tw[i] >> 24) / 255.0f);
        }
    } else {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            h[i] = p.penumbra.load(px0 + i, py);
            z[i] = p.viewZ.load(px0 + i, py);
            if (TR) st[i] = p.translucency.load(px0 + i, py);
        }
    }
    int nLit = 0, nUmbra = 0, nInf = 0;
    float radius = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const float viewZ = sigmaUnpackViewZ(cb, z[i]);
        const bool isInf = !sigmaInRange(cb, viewZ), isShadow = h[i] == 0.0f, isLit = sigmaIsLit(h[i]);
        bool isOpaque = true;
        if (TR) isOpaque = dot(make_float3(st[i].y, st[i].z, st[i].w), make_float3(0.2126f, 0.7152f, 0.0722f)) < 0.003f;   // SIGMA_ClassifyTiles.cs.hlsl:55-58
        nLit += (isLit || isInf || isShadow) ? 1 : 0;
        nUmbra += ((!isLit && isOpaque) || isInf || isShadow) ? 1 : 0;
        nInf += isInf ? 1 : 0;
        const float hitDist = (isLit || isInf) ? 0.0f : h[i];
        const float pixelSize = pixelRadiusToWorld(cb.unproject, cb.orthoMode, 1.0f, viewZ);
        radius = fmaxf(radius, sigmaKernelRadiusInPixels(hitDist, pixelSize));   // radii are >= 0 ( InterlockedMax on asuint in the reference ); NaN drops out like there
    }
    // one packed add for the three counters ( each <= 256: 10 bits apiece ), one max for the radius
    const uint32_t counts = __reduce_add_sync(0xFFFFFFFFu, (uint32_t)nLit | ((uint32_t)nUmbra << 10) | ((uint32_t)nInf << 20));
    const uint32_t m = __reduce_max_sync(0xFFFFFFFFu, __float_as_uint(fmaxf(radius, 0.0f)));
    if (lane == 0) {
        const bool lit = (counts & 1023u) == 256u, umbra = ((counts >> 10) & 1023u) == 256u, inf = (counts >> 20) == 256u;
        p.outTiles.store(tx, ty, make_float4((lit || umbra) ? 0.0f : 1.0f, saturate(__uint_as_float(m) / 16.0f), inf ? 1.0f : 0.0f, 0.0f));
    }
}

// One thread per tile texel
__global__ void __launch_bounds__(256) sigmaSmoothTilesKernel(const __grid_constant__ SigmaConstants cb, const __grid_constant__ SigmaSmoothTilesParams p) {
    pdlEntry();
    const int x = blockIdx.x * 16 + (threadIdx.x & 15), y = blockIdx.y * 16 + (threadIdx.x >> 4);
    const float4 center = p.tiles.load(x, y);
    const float k = 1.01f / (center.y + 0.01f);
    float blurry = 0.0f, sum = 0.0f;
#pragma unroll
    for (int j = 0; j <= 2; j++)
#pragma unroll
        for (int i = 0; i <= 2; i++) {
            const float d2 = (float)((i - 1) * (i - 1) + (j - 1) * (j - 1));  // length( float2( i, j ) - 1 ) ^ 2
            const float w = exp2f(-k * d2);
            blurry += p.tiles.load(clampi(x + i - 1, 0, cb.tilesSizeMinusOne[0]), clampi(y + j - 1, 0, cb.tilesSizeMinusOne[1])).x * w;
            sum += w;
        }
    p.outTiles.store(x, y, make_float2(center.z, blurry / sum));
}

template <class ST>
__global__ void __launch_bounds__(BLOCK_W* BLOCK_H) sigmaCopyKernel(const __grid_constant__ SigmaConstants cb, const __grid_constant__ SigmaCopyParams<ST> p, int ctaY0) {
    pdlEntry();
    using Raw = typename std::conditional<std::is_same<ST, TexR8>::value, uint8_t, uint32_t>::type;
    const int px = blockIdx.x * BLOCK_W + threadIdx.x, py = (blockIdx.y + ctaY0) * BLOCK_H + threadIdx.y;
    // the reference dispatches this pass over the PREVIOUS frame's rect ( Sigma_Shadow.hpp: "USE_PREV_DIMS" ) in 8x16 groups
    if (px >= (((int)cb.rectSizePrev[0] + 7) & ~7) || py >= (((int)cb.rectSizePrev[1] + 15) & ~15)) return;
    const float isSky = p.tiles.load(px >> 4, py >> 4).x;
    if ((isSky != 0.0f && !cb.isRectChanged) || !p.history.inside(px, py)) return;
    *p.outHistory.template ptrw<Raw>(px, py) = __ldg(p.history.template ptr<Raw>(px, py));
    p.outHistoryLength.store(px, py, p.historyLength.load(px, py));
}

// ---------------------------------------------------------------------------------------------------------------
// Both blur passes ( SIGMA_Blur.cs.hlsl:48-286 ): a dense 5x5 that estimates the penumbra size and pre-filters, then 8 rotated sparse taps at that radius.
// What a tap contributes that does not depend on the pixel it is a tap OF is decided once per texel while the tile is staged: "counts at all" ( in
// denoising range and not umbra-vs-penumbra mismatched with a centre that got this far, i.e. penumbra != 0 ) and "is in penumbra" ( not lit ) go into shared
// memory next to { penumbra, viewZ } as 0 / 1 factors, so the 24 dense taps are branch- and select-free multiply-adds. dot( Nv, Xv( tap ) ) is affine in the
// tap's uv: the per-column and per-row terms are hoisted out of the loop. ( The sums are the reference's, regrouped: differences are rounding-level. )
#ifndef SIGMA_BLUR_MIN_BLOCKS
#define SIGMA_BLUR_MIN_BLOCKS 5   // <= 51 registers: five CTAs per SM. The first pass ( with the folded copy ) sat at 54 registers = 4 CTAs, 46 % occupancy, top stall long_scoreboard
#endif
template <bool FIRST_PASS, bool TR>
__global__ void __launch_bounds__(BLOCK_W* BLOCK_H, SIGMA_BLUR_MIN_BLOCKS) sigmaBlurKernel(const __grid_constant__ SigmaConstants cb,
                                                                   const __grid_constant__ SigmaBlurParams<typename SigmaSignal<TR>::Tex> p, int ctaY0) {
    pdlEntry();
    using SG = SigmaSignal<TR>;
    using S = typename SG::T;
    constexpr bool SHADOW_FROM_PENUMBRA = FIRST_PASS && !TR;  // s = IsLit( penumbra ): no shadow texture bound (SIGMA_Blur.cs.hlsl:35-39)
    __shared__ float4 sTap[TILE_H][TILE_W];   // { penumbra ( <= FP16_MAX ), viewZ, counts ? 1 : 0, lit ? 0 : 1 }
    __shared__ S sShadow[SHADOW_FROM_PENUMBRA ? 1 : TILE_H][SHADOW_FROM_PENUMBRA ? 1 : TILE_W];

    // CTA order: first pass default, post-blur reversed (SIGMA_Blur.cs.hlsl:51-55)
    const int bx = FIRST_PASS ? (int)blockIdx.x : (int)(gridDim.x - 1u - blockIdx.x), by = FIRST_PASS ? (int)(blockIdx.y + ctaY0) : (int)(gridDim.y - 1u - blockIdx.y) + ctaY0;
    const int px = bx * BLOCK_W + threadIdx.x, py = by * BLOCK_H + threadIdx.y;

    // The CTA covers two 16x16 tiles of one tile row
    const float skyL = p.tiles.load((bx * BLOCK_W) >> 4, py >> 4).x, skyR = p.tiles.load((bx * BLOCK_W + 16) >> 4, py >> 4).x;
    if (skyL != 0.0f && skyR != 0.0f) return;

    if (FIRST_PASS && p.copy) {
        // SIGMA_Copy.cs.hlsl:20-34 for this thread's pixel ( same grid: the executor only folds the pass in when the rect is the whole, unchanged frame ): two loads
        // and two stores in flight under the tile staging below, instead of a launch of their own
        using Raw = typename std::conditional<TR, uint32_t, uint8_t>::type;
        if ((threadIdx.x < 16 ? skyL : skyR) == 0.0f && p.copyHistory.inside(px, py)) {
            *p.copyOutHistory.template ptrw<Raw>(px, py) = __ldg(p.copyHistory.template ptr<Raw>(px, py));
            p.copyOutHistoryLength.store(px, py, p.copyHistoryLength.load(px, py));
        }
    }

    {
        const int baseX = bx * BLOCK_W - SIGMA_BORDER, baseY = by * BLOCK_H - SIGMA_BORDER;
        const int tid = threadIdx.y * BLOCK_W + threadIdx.x;
        for (int i = tid; i < TILE_W * TILE_H; i += BLOCK_W * BLOCK_H) {
            const int sx = i % TILE_W, sy = i / TILE_W;
            const int gx = clampi(baseX + sx, 0, cb.rectSizeMinusOne[0]), gy = clampi(baseY + sy, 0, cb.rectSizeMinusOne[1]);
            const float pen = p.penumbra.load(gx, gy), z = sigmaUnpackViewZ(cb, p.viewZ.load(gx, gy));
            sTap[sy][sx] = make_float4(fminf(pen, NRD_FP16_MAX), z, (pen != 0.0f && sigmaInRange(cb, z)) ? 1.0f : 0.0f, sigmaIsLit(pen) ? 0.0f : 1.0f);
            if (!SHADOW_FROM_PENUMBRA) {
                const S s = p.shadow.load(gx, gy);
                sShadow[SHADOW_FROM_PENUMBRA ? 0 : sy][SHADOW_FROM_PENUMBRA ? 0 : sx] = FIRST_PASS ? s : s * s;  // SIGMA_BackEnd_UnpackShadow
            }
        }
    }
    __syncthreads();

    const float isSky = threadIdx.x < 16 ? skyL : skyR;
    if (isSky != 0.0f || px > cb.rectSizeMinusOne[0] || py > cb.rectSizeMinusOne[1]) return;

    const int smx = threadIdx.x + SIGMA_BORDER, smy = threadIdx.y + SIGMA_BORDER;
    const float4 centerData = sTap[smy][smx];
    const float viewZ = centerData.y;
    if (!sigmaInRange(cb, viewZ)) return;
    const float centerPenumbra = p.penumbra.load(px, py);   // as stored ( the staged copy is clamped to FP16_MAX ); an L1 hit: the tile load just read it

    const float2 rectSizeInv = make_float2(cb.rectSizeInv[0], cb.rectSizeInv[1]);
    const float2 pixelUv = make_float2((float)px + 0.5f, (float)py + 0.5f) * rectSizeInv;
    const float tileValue = sigmaTileValue(p.tiles, pixelUv * make_float2(cb.resolutionScale[0], cb.resolutionScale[1]));

    // a tap's shadow: IsLit( penumbra ) in the first pass of SIGMA_SHADOW ( = 1 - the "in penumbra" factor ), the staged texel otherwise
    auto shadowAt = [&](int y, int x, float inPenumbra) -> S {
        if (SHADOW_FROM_PENUMBRA) return SG::splat(1.0f - inPenumbra);
        return sShadow[SHADOW_FROM_PENUMBRA ? 0 : y][SHADOW_FROM_PENUMBRA ? 0 : x];
    };

    if (tileValue == 0.0f || centerPenumbra == 0.0f) {
        if (FIRST_PASS || cb.stabilizationStrength != 0.0f) p.outPenumbra.store(px, py, centerPenumbra);
        p.outShadow.store(px, py, sigmaPackShadow(shadowAt(smy, smx, centerData.w)));
        return;
    }

    const float3 Xv = reconstructViewPosition(pixelUv, cb.frustum, viewZ, cb.orthoMode);
    const float3 N = xyz(unpackNormalRoughness(p.normalRoughness.loadRaw(px, py)));
    const float3 Nv = rotate(cb.worldToView, N);

    const float pixelSize = pixelRadiusToWorld(cb.unproject, cb.orthoMode, 1.0f, viewZ);
    const float invPixelSize = 1.0f / pixelSize;
    const float frustumSize = frustumSizeAt(cb.minRectDimMulUnproject, cb.orthoMode, viewZ);
    const float3 Vv = cb.orthoMode == 0.0f ? normalize(-Xv) : make_float3(0.0f, 0.0f, -1.0f);
    const float NoV = fabsf(dot(Nv, Vv));
    const float2 geomParams = geometryWeightParams(cb.planeDistSensitivity, frustumSize, Xv, Nv);

    // dot( Nv, Xv( uv, z ) ) = ( Nv.x * ( u * f2 + f0 ) + Nv.y * ( v * f3 + f1 ) ) * ( ortho ? ortho : z ) + Nv.z * z = ( u * nu + v * nv + n0 ) * scale + Nv.z * z
    const float nu = Nv.x * cb.frustum[2], nv = Nv.y * cb.frustum[3], n0 = Nv.x * cb.frustum[0] + Nv.y * cb.frustum[1];
    const bool perspective = cb.orthoMode == 0.0f;
    // ( the geometry weight's affine map folded in: x * a + b with x = NoX )
    const float ga = geomParams.x, gb = geomParams.y;
    auto geometryWeight = [&](float planeTerm, float z) {   // planeTerm = u * nu + v * nv + n0
        const float NoX = planeTerm * (perspective ? z : cb.orthoMode) + Nv.z * z;
        const float t = satOneMinusAbs(NoX * ga + gb);   // one FADD.SAT ( pairmath.cuh )
        return (t * t) * fmaf(t, -2.0f, 3.0f);
    };

    // Estimate penumbra size and filter shadow ( dense 5x5 )
    float colTerm[SIGMA_BORDER * 2 + 1], rowTerm[SIGMA_BORDER * 2 + 1];
#pragma unroll
    for (int i = 0; i <= SIGMA_BORDER * 2; i++) {
        colTerm[i] = (pixelUv.x + (float)(i - SIGMA_BORDER) * rectSizeInv.x) * nu;
        rowTerm[i] = (pixelUv.y + (float)(i - SIGMA_BORDER) * rectSizeInv.y) * nv + n0;
    }
    float sumX = 1.0f;
    S centerTap = shadowAt(smy, smx, centerData.w), result = centerTap;
    // the centre: weight 1, no geometry test
    float sumY = centerData.w / (1.0f + centerData.x * invPixelSize);
    float penumbra = centerData.x * sumY;
#pragma unroll
    for (int j = 0; j <= SIGMA_BORDER * 2; j++)
#pragma unroll
        for (int i = 0; i <= SIGMA_BORDER * 2; i++) {
            if (i == SIGMA_BORDER && j == SIGMA_BORDER) continue;
            const float4 tap = sTap[threadIdx.y + j][threadIdx.x + i];   // { penumbra, viewZ, counts, inPenumbra }
            const S s = shadowAt(threadIdx.y + j, threadIdx.x + i, tap.w);
            const float g = kGaussian5x5[(i - SIGMA_BORDER) * (i - SIGMA_BORDER) + (j - SIGMA_BORDER) * (j - SIGMA_BORDER)];
            float w = (g * tap.z) * geometryWeight(colTerm[i] + rowTerm[j], tap.y);
            result += s * w;
            sumX += w;
            w *= tap.w / (1.0f + tap.x * invPixelSize);   // pixelSize / ( pixelSize + penumbra ), 0 where lit
            penumbra += tap.x * w;
            sumY += w;
        }
    result /= sumX;
    sumX = 1.0f;
    penumbra /= fmaxf(sumY, NRD_EPS);
    sumY = sumY != 0.0f ? 1.0f : 0.0f;

    // Avoid blurry result if penumbra size < NRD_BORDER
    const float penumbraInPixels = penumbra * invPixelSize;
    float f = smoothStep(0.0f, (float)SIGMA_BORDER, penumbraInPixels);
    result = lerp(centerTap, result, f);

    // Sparse pass: avoid unnecessary weight increase for the unfiltered center sample if the blur radius is small
    f = lerp(4.0f, 1.0f, f);
    result *= f;
    penumbra *= f;
    sumX *= f;
    sumY *= f;

    const float blurRadius = sigmaKernelRadiusInPixels(penumbra, pixelSize, tileValue);
    const float* rot = FIRST_PASS ? cb.rotator : cb.rotatorPost;  // SIGMA_ROTATOR_MODE = NRD_FRAME
    const float4 rotator = make_float4(rot[0], rot[1], rot[2], rot[3]);

    float2 skew = lerp(f2(1.0f) - fabs2(make_float2(Nv.x, Nv.y)), f2(1.0f), NoV);
    skew = skew / fmaxf(skew.x, skew.y);
    skew = skew * (rectSizeInv * blurRadius);
    const float4 scaledRotator = scaleRotator(rotator, skew);

    const float invEstimatedPenumbra = 1.0f / fmaxf(penumbra, NRD_EPS);
    const float2 rectSize = make_float2(cb.rectSize[0], cb.rectSize[1]);
    const int rectW = cb.rectSizeMinusOne[0] + 1, rectH = cb.rectSizeMinusOne[1] + 1;
#pragma unroll
    for (int n = 0; n < 8; n++) {
        const float2 uvTap = pixelUv + rotate2(scaledRotator, make_float2(kSpecial8[n][0], kSpecial8[n][1]));
        // snap to the pixel center: the texel index is floor( uv * rectSize ) for every texture of the pass — point-sampling the snapped uv scaled by gResolutionScale
        // and clamped to the viewport ( ClampUvToViewport, Common.hlsli:242 ) lands on exactly that texel, clamped to the rect
        const int tx = __float2int_rd(uvTap.x * rectSize.x), ty = __float2int_rd(uvTap.y * rectSize.y);
        const float2 uv = make_float2(((float)tx + 0.5f) * rectSizeInv.x, ((float)ty + 0.5f) * rectSizeInv.y);
        const bool inScreen = (unsigned)tx < (unsigned)rectW && (unsigned)ty < (unsigned)rectH;   // IsInScreenNearest( uv ): 0 < ( k + 0.5 ) / size < 1
        const int cx = clampi(tx, 0, rectW - 1), cy = clampi(ty, 0, rectH - 1);

        const float penumRaw = p.penumbra.fetch(cx, cy);
        const float penum = fminf(penumRaw, NRD_FP16_MAX);   // finite: the products below need no "w == 0" guards
        const float zs = sigmaUnpackViewZ(cb, p.viewZ.fetch(cx, cy));
        const bool lit = sigmaIsLit(penumRaw);
        S s;
        if (SHADOW_FROM_PENUMBRA)
            s = SG::splat(lit ? 1.0f : 0.0f);
        else {
            s = p.shadow.fetch(cx, cy);
            if (!FIRST_PASS) s *= s;
        }

        float w = (inScreen && penum != 0.0f && sigmaInRange(cb, zs)) ? kGaussian5x5[(int)(kSpecial8[n][2] * kSpecial8[n][2] * 4.0f)] : 0.0f;   // ( the centre's penumbra is not 0 here )
        w *= saturate(penum * invEstimatedPenumbra);  // avoid umbra leaking inside wide penumbra
        w *= geometryWeight(uv.x * nu + (uv.y * nv + n0), zs);

        result += s * w;
        sumX += w;
        w = lit ? 0.0f : w / (1.0f + penum * invPixelSize);   // x pixelSize / ( pixelSize + penumbra )
        penumbra += penum * w;
        sumY += w;
    }

    result /= sumX;
    penumbra = sumY == 0.0f ? centerPenumbra : penumbra / sumY;

    if (FIRST_PASS || cb.stabilizationStrength != 0.0f) p.outPenumbra.store(px, py, penumbra);
    p.outShadow.store(px, py, sigmaPackShadow(result));
}

// ---------------------------------------------------------------------------------------------------------------
NRD_DEV uint32_t packViewZAndHistoryLength(float viewZ, float historyLength) {
    return (__float_as_uint(viewZ) & ~7u) | min((uint32_t)(historyLength + 0.5f), 7u);
}

template <bool TR>
__global__ void __launch_bounds__(BLOCK_W* BLOCK_H)
    sigmaTemporalStabilizationKernel(const __grid_constant__ SigmaConstants cb, const __grid_constant__ SigmaTemporalStabilizationParams<typename SigmaSignal<TR>::Tex> p, int ctaY0) {
    pdlEntry();
    using SG = SigmaSignal<TR>;
    using S = typename SG::T;
    __shared__ S sShadow[TILE_H][TILE_W];
    __shared__ float sPenumbra[TILE_H][TILE_W];

    const int bx = blockIdx.x, by = blockIdx.y + ctaY0;  // NRD_CTA_ORDER_DEFAULT
    const int px = bx * BLOCK_W + threadIdx.x, py = by * BLOCK_H + threadIdx.y;
    const float skyL = p.tiles.load((bx * BLOCK_W) >> 4, py >> 4).x, skyR = p.tiles.load((bx * BLOCK_W + 16) >> 4, py >> 4).x;
    if (skyL != 0.0f && skyR != 0.0f) return;
    {
        const int baseX = bx * BLOCK_W - SIGMA_BORDER, baseY = by * BLOCK_H - SIGMA_BORDER;
        const int tid = threadIdx.y * BLOCK_W + threadIdx.x;
        for (int i = tid; i < TILE_W * TILE_H; i += BLOCK_W * BLOCK_H) {
            const int sx = i % TILE_W, sy = i / TILE_W;
            const int gx = clampi(baseX + sx, 0, cb.rectSizeMinusOne[0]), gy = clampi(baseY + sy, 0, cb.rectSizeMinusOne[1]);
            const S s = p.shadow.load(gx, gy);
            sShadow[sy][sx] = s * s;
            sPenumbra[sy][sx] = p.penumbra.load(gx, gy);
        }
    }
    __syncthreads();

    const float isSky = threadIdx.x < 16 ? skyL : skyR;
    const float viewZ = sigmaUnpackViewZ(cb, p.viewZ.load(px, py));
    if (isSky != 0.0f || px > cb.rectSizeMinusOne[0] || py > cb.rectSizeMinusOne[1] || !sigmaInRange(cb, viewZ)) return;

    const int smx = threadIdx.x + SIGMA_BORDER, smy = threadIdx.y + SIGMA_BORDER;
    const float centerPenumbra = sPenumbra[smy][smx];
    const float2 pixelUv = make_float2((float)px + 0.5f, (float)py + 0.5f) * make_float2(cb.rectSizeInv[0], cb.rectSizeInv[1]);
    const float tileValue = sigmaTileValue(p.tiles, pixelUv * make_float2(cb.resolutionScale[0], cb.resolutionScale[1]));

    if (tileValue == 0.0f || centerPenumbra == 0.0f) {  // hard shadow / fully lit: nothing to stabilise
        p.outShadow.store(px, py, sigmaPackShadow(sShadow[smy][smx]));
        p.outHistoryLength.store(px, py, packViewZAndHistoryLength(viewZ, SIGMA_MAX_ACCUM_FRAME_NUM));
        return;
    }

    // Local variance
    float sum = 0.0f;
    S m1 = SG::splat(0.0f), m2 = SG::splat(0.0f), input = SG::splat(0.0f);
#pragma unroll
    for (int j = 0; j <= SIGMA_BORDER * 2; j++)
#pragma unroll
        for (int i = 0; i <= SIGMA_BORDER * 2; i++) {
            const S s = sShadow[threadIdx.y + j][threadIdx.x + i];
            float w = 1.0f;
            if (i == SIGMA_BORDER && j == SIGMA_BORDER)
                input = s;
            else {
                const float penum = sPenumbra[threadIdx.y + j][threadIdx.x + i];
                w = sigmaBothLitOrUnlit(centerPenumbra, penum);
                w *= gaussianWeight(length(make_float2((float)(i - SIGMA_BORDER), (float)(j - SIGMA_BORDER)) / (float)SIGMA_BORDER));
            }
            m1 += s * w;
            m2 += s * s * w;
            sum += w;
        }
    m1 /= sum;
    m2 /= sum;
    S sigma = SG::stdDev(m1, m2);

    // Current and previous positions
    const float3 Xv = reconstructViewPosition(pixelUv, cb.frustum, viewZ, cb.orthoMode);
    const float3 X = rotateInverse(cb.worldToView, Xv);
    const float4 mvRaw = p.mv.load(px, py);
    float3 mv = make_float3(mvRaw.x * cb.mvScale[0], mvRaw.y * cb.mvScale[1], mvRaw.z * cb.mvScale[2]);
    float3 Xprev = X;
    float2 smbPixelUv = pixelUv + make_float2(mv.x, mv.y);
    if (cb.mvScale[3] == 0.0f) {
        if (cb.mvScale[2] == 0.0f) mv.z = affine(cb.worldToViewPrev, X).z - viewZ;
        const float viewZprev = viewZ + mv.z;
        const float3 Xvprevlocal = reconstructViewPosition(smbPixelUv, cb.frustumPrev, viewZprev, cb.orthoMode);
        Xprev = rotateInverse(cb.worldToViewPrev, Xvprevlocal) + make_float3(cb.cameraDelta[0], cb.cameraDelta[1], cb.cameraDelta[2]);
    } else {
        Xprev = Xprev + mv;
        smbPixelUv = screenUv(cb.worldToClipPrev, Xprev);
    }

    // History length: 2x2 footprint of { viewZ with the length in the low 3 bits }
    const float2 rectSizePrev = make_float2(cb.rectSizePrev[0], cb.rectSizePrev[1]);
    const Bilinear smbFilter = getBilinearFilter(smbPixelUv, rectSizePrev);
    const int ox = (int)smbFilter.origin.x, oy = (int)smbFilter.origin.y;
    const uint32_t d00 = p.historyLength.fetchClamped(ox, oy), d10 = p.historyLength.fetchClamped(ox + 1, oy);
    const uint32_t d01 = p.historyLength.fetchClamped(ox, oy + 1), d11 = p.historyLength.fetchClamped(ox + 1, oy + 1);
    const float4 prevViewZ = make_float4(__uint_as_float(d00 & ~7u), __uint_as_float(d10 & ~7u), __uint_as_float(d01 & ~7u), __uint_as_float(d11 & ~7u));
    const float4 prevHistoryLength = make_float4((float)(d00 & 7u), (float)(d10 & 7u), (float)(d01 & 7u), (float)(d11 & 7u));

    const float frustumSize = frustumSizeAt(cb.minRectDimMulUnproject, cb.orthoMode, viewZ);
    float4 disocclusionThreshold = f4(disocclusionThresholdAt(SIGMA_DISOCCLUSION_THRESHOLD, frustumSize, 1.0f));
    disocclusionThreshold = disocclusionThreshold * isInScreenBilinear(smbFilter.origin, rectSizePrev);
    disocclusionThreshold = disocclusionThreshold - NRD_EPS;

    const float3 Xvprev = affine(cb.worldToViewPrev, Xprev);
    const float4 planeDist = make_float4(fabsf(prevViewZ.x - Xvprev.z), fabsf(prevViewZ.y - Xvprev.z), fabsf(prevViewZ.z - Xvprev.z), fabsf(prevViewZ.w - Xvprev.z));
    // step( planeDist, threshold ): 1 where threshold >= planeDist
    const float4 occlusion = make_float4(disocclusionThreshold.x >= planeDist.x ? 1.0f : 0.0f, disocclusionThreshold.y >= planeDist.y ? 1.0f : 0.0f,
                                         disocclusionThreshold.z >= planeDist.z ? 1.0f : 0.0f, disocclusionThreshold.w >= planeDist.w ? 1.0f : 0.0f);
    const float4 occlusionWeights = bilinearCustomWeights(smbFilter, occlusion);
    float historyLength = applyCustomWeights(prevHistoryLength.x, prevHistoryLength.y, prevHistoryLength.z, prevHistoryLength.w, occlusionWeights);

    // Sample history
    const bool isCatRomAllowed = sum4(occlusionWeights) > 3.5f;
    const HistoryFilter hf(saturate(smbPixelUv) * rectSizePrev, make_float2(cb.resourceSizeInvPrev[0], cb.resourceSizeInvPrev[1]), occlusionWeights, isCatRomAllowed);
    S history = hf.color(p.history);
    history = saturate(history);
    history *= history;

    // Clamp history
    sigma *= lerp(SIGMA_TS_SIGMA_SCALE, 1.0f, 1.0f / (1.0f + historyLength));
    S historyClamped = SG::clamp(history, m1 - sigma, m1 + sigma);

    // Antilag ( on the shadow channel: SIGMA_TemporalStabilization.cs.hlsl:183 )
    float antilag = sqrt01(fabsf(SG::x(historyClamped) - SG::x(history)));
    antilag = saturate(1.0f - antilag);
    historyLength *= antilag;

    const float historyWeight = historyLength / (1.0f + historyLength);
    const float streetMagic = 0.6f * historyWeight * antilag;
    historyClamped = lerp(historyClamped, history, streetMagic);

    const S result = lerp(input, historyClamped, fminf(cb.stabilizationStrength, historyWeight));
    historyLength = fminf(historyLength + 1.0f, SIGMA_MAX_ACCUM_FRAME_NUM);

    p.outShadow.store(px, py, sigmaPackShadow(result));
    p.outHistoryLength.store(px, py, packViewZAndHistoryLength(viewZ, historyLength));
}

template <bool TR>
__global__ void __launch_bounds__(BLOCK_W* BLOCK_H) sigmaSplitScreenKernel(const __grid_constant__ SigmaConstants cb,
                                                                          const __grid_constant__ SigmaSplitScreenParams<typename SigmaSignal<TR>::Tex> p, int ctaY0) {
    pdlEntry();
    using SG = SigmaSignal<TR>;
    const int px = blockIdx.x * BLOCK_W + threadIdx.x, py = (blockIdx.y + ctaY0) * BLOCK_H + threadIdx.y;
    const float u = ((float)px + 0.5f) * cb.rectSizeInv[0];
    if (u > cb.splitScreen || px > cb.rectSizeMinusOne[0] || py > cb.rectSizeMinusOne[1]) return;
    const float viewZ = sigmaUnpackViewZ(cb, p.viewZ.load(px, py));
    typename SG::T s;
    if (TR)
        s = p.translucency.load(px, py);
    else
        s = SG::splat(sigmaIsLit(p.penumbra.load(px, py)) ? 1.0f : 0.0f);
    p.outShadow.store(px, py, s * (sigmaInRange(cb, viewZ) ? 1.0f : 0.0f));
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// Dispatch by shader identifier (called by the executor). Returns an nrd::Result; `err` explains a failure.
// ---------------------------------------------------------------------------------------------------------------
namespace {
uint32_t bytesOf(nrd::Format f) {
    switch (f) {
        case nrd::Format::R8_UNORM: return 1;
        case nrd::Format::RG8_UNORM: case nrd::Format::R16_SFLOAT: return 2;
        case nrd::Format::RGBA16_SFLOAT: return 8;
        default: return 4;
    }
}
struct SigmaBinder {
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
};
}  // namespace

// The Copy pass of a frame denoised through a context ( sigmaSetCopyFusion, executor.cu ) is not launched when it arrives: its bindings wait here for the first
// blur pass, which does the copy for its own pixels. Anything else arriving first, or the end of the frame, launches it on its own.
struct PendingCopy {
    bool armed = false, wide = false;
    Rows rows;
    SigmaConstants cb;
    SigmaCopyParams<TexR8> narrowParams;
    SigmaCopyParams<TexRGBA8> wideParams;
};
thread_local PendingCopy g_pendingCopy;
thread_local bool g_fuseCopy = false;

static void launchCopy(const SigmaConstants& cb, bool wide, const SigmaCopyParams<TexR8>& pn, const SigmaCopyParams<TexRGBA8>& pw, Rows rows, cudaStream_t stream) {
    const dim3 block(BLOCK_W, BLOCK_H);
    const RowGrid g = rowGrid(rows, ((int)cb.rectSizePrev[1] + 15) / 16 * 16, BLOCK_H);   // the previous rect, rounded to the reference's 8x16 groups
    if (!g.count) return;
    const dim3 prevGrid(((int)cb.rectSizePrev[0] + BLOCK_W - 1) / BLOCK_W, g.count);
    if (wide) launchK(sigmaCopyKernel<TexRGBA8>, prevGrid, block, 0, stream, cb, pw, g.ctaY0);
    else launchK(sigmaCopyKernel<TexR8>, prevGrid, block, 0, stream, cb, pn, g.ctaY0);
}
void sigmaFlushPendingCopy(cudaStream_t stream) {
    if (!g_pendingCopy.armed) return;
    g_pendingCopy.armed = false;
    launchCopy(g_pendingCopy.cb, g_pendingCopy.wide, g_pendingCopy.narrowParams, g_pendingCopy.wideParams, g_pendingCopy.rows, stream);
}
void sigmaSetCopyFusion(bool on) {
    g_fuseCopy = on;
    if (!on) g_pendingCopy.armed = false;   // a frame that ended in an error leaves nothing behind
}

uint32_t dispatchSigma(const PipelineKey& key, const void* constants, uint32_t constantsSize, const nrdcuTexture* tex, uint32_t n, Rows rows, cudaStream_t stream, std::string& err) {
    using nrd::Format;
    using nrd::Result;
    const char* const id = key.id;
    if (constantsSize != sizeof(SigmaConstants) || !constants) {
        err = std::string(id) + ": expected " + std::to_string(sizeof(SigmaConstants)) + " constant bytes";
        return (uint32_t)Result::INVALID_ARGUMENT;
    }
    SigmaConstants cb;
    memcpy(&cb, constants, sizeof(cb));
    if (cb.rectOrigin[0] || cb.rectOrigin[1]) {   // NRD_SUPPORTS_VIEWPORT_OFFSET = 0; dynamic resolution ( rectSize < resourceSize ) itself is supported
        err = std::string(id) + ": rectOrigin must be 0";
        return (uint32_t)Result::INVALID_ARGUMENT;
    }
    SigmaBinder b{tex, n, 0, true, &err, id};
    auto bad = [&](uint32_t expected) {
        if (b.ok && b.next == expected && n == expected) return false;
        if (err.empty()) err = std::string(id) + ": wrong number of textures";
        return true;
    };
    const dim3 block(BLOCK_W, BLOCK_H);
    // rows [ begin, end ) of the rect ( multi-GPU strips, nrdcuDenoiseRows ): the per-pixel passes cover only the CTA rows of the strip; the two tile passes
    // always run over the whole frame ( every strip holds the full-frame inputs they read, and the bicubic tile lookup of the blur passes reaches two tile
    // rows past the strip: 12 us of redundant work per frame instead of a seam trade of the tile masks )
    const RowGrid rg = rowGrid(rows, cb.rectSizeMinusOne[1] + 1, BLOCK_H);
    const int ctaY0 = rg.ctaY0;
    const dim3 pixelGrid((cb.rectSizeMinusOne[0] + BLOCK_W) / BLOCK_W, rg.count);
    const bool wholeFrame = rows.begin <= 0 && rows.end > cb.rectSizeMinusOne[1];
    const int tilesW = cb.tilesSizeMinusOne[0] + 1, tilesH = cb.tilesSizeMinusOne[1] + 1;
    if (rg.count == 0 && key.pass != SIGMA_CLASSIFY_TILES && key.pass != SIGMA_SMOOTH_TILES && key.pass != SIGMA_COPY) return (uint32_t)Result::SUCCESS;

    const bool tr = key.translucency;
    if (!(key.pass == SIGMA_BLUR && key.firstPass) && key.pass != SIGMA_COPY) sigmaFlushPendingCopy(stream);
    if (key.pass == SIGMA_CLASSIFY_TILES) {
        SigmaClassifyTilesParams p = {};
        p.viewZ = b.take<TexR32F>(Format::R32_SFLOAT);
        p.penumbra = b.take<TexR16F>(Format::R16_SFLOAT);
        if (tr) p.translucency = b.take<TexRGBA8>(Format::RGBA8_UNORM);
        p.outTiles = b.take<TexRGBA8>(Format::RGBA8_UNORM);
        if (bad(tr ? 4 : 3)) return (uint32_t)Result::INVALID_ARGUMENT;
        // 128-bit row loads need 16-byte aligned rows ( pool textures always are; the application's IN_VIEWZ / IN_PENUMBRA / IN_TRANSLUCENCY usually )
        auto aligned = [](const TexView& t, uint32_t texelBytes) { return ((uintptr_t)t.data % 16u) == 0 && ((size_t)t.pitch * texelBytes) % 16u == 0; };
        const bool vec = aligned(p.viewZ, 4) && aligned(p.penumbra, 2) && (!tr || aligned(p.translucency, 4));
        const dim3 grid((tilesW + 7) / 8, tilesH);
        if (tr) {
            if (vec) launchK(sigmaClassifyTilesKernel<true, true>, grid, 256, 0, stream, cb, p, tilesW, tilesH);
            else launchK(sigmaClassifyTilesKernel<true, false>, grid, 256, 0, stream, cb, p, tilesW, tilesH);
        } else {
            if (vec) launchK(sigmaClassifyTilesKernel<false, true>, grid, 256, 0, stream, cb, p, tilesW, tilesH);
            else launchK(sigmaClassifyTilesKernel<false, false>, grid, 256, 0, stream, cb, p, tilesW, tilesH);
        }
    } else if (key.pass == SIGMA_SMOOTH_TILES) {
        SigmaSmoothTilesParams p;
        p.tiles = b.take<TexRGBA8>(Format::RGBA8_UNORM);
        p.outTiles = b.take<TexRG8>(Format::RG8_UNORM);
        if (bad(2)) return (uint32_t)Result::INVALID_ARGUMENT;
        launchK(sigmaSmoothTilesKernel, dim3((tilesW + 15) / 16, (tilesH + 15) / 16), 256, 0, stream, cb, p);
    } else if (key.pass == SIGMA_COPY) {
        // one shader for both variants (gIn_History is declared float4): the history's own format picks the texel width
        const bool wide = n > 1 && tex[1].format == (uint32_t)Format::RGBA8_UNORM;
        auto run = [&](auto sig) -> bool {
            using SG = decltype(sig);
            SigmaCopyParams<typename SG::Tex> p;
            p.tiles = b.take<TexRG8>(Format::RG8_UNORM);
            p.history = b.take<typename SG::Tex>(SG::format);
            p.historyLength = b.take<TexR32U>(Format::R32_UINT);
            p.outHistory = b.take<typename SG::Tex>(SG::format);
            p.outHistoryLength = b.take<TexR32U>(Format::R32_UINT);
            if (bad(5)) return false;
            // folded into the first blur pass when both cover the same pixels: the rect is the whole texture and did not change ( otherwise this pass runs over
            // the previous rect, rounded to the reference's 8x16 groups )
            const bool sameGrid = !cb.isRectChanged && (int)cb.rectSizePrev[0] == p.history.w && (int)cb.rectSizePrev[1] == p.history.h && cb.rectSizeMinusOne[0] + 1 == p.history.w &&
                                  cb.rectSizeMinusOne[1] + 1 == p.history.h;
            g_pendingCopy.cb = cb;
            g_pendingCopy.rows = rows;
            g_pendingCopy.wide = std::is_same<SG, SigmaSignal<true>>::value;
            if constexpr (std::is_same<SG, SigmaSignal<true>>::value) g_pendingCopy.wideParams = p;
            else g_pendingCopy.narrowParams = p;
            g_pendingCopy.armed = true;
            if (!(g_fuseCopy && sameGrid && wholeFrame)) sigmaFlushPendingCopy(stream);
            return true;
        };
        if (!(wide ? run(SigmaSignal<true>()) : run(SigmaSignal<false>()))) return (uint32_t)Result::INVALID_ARGUMENT;
    } else if (key.pass == SIGMA_BLUR) {
        const bool first = key.firstPass;
        auto run = [&](auto sig) -> bool {
            using SG = decltype(sig);
            constexpr bool TR = std::is_same<SG, SigmaSignal<true>>::value;
            SigmaBlurParams<typename SG::Tex> p = {};
            p.viewZ = b.take<TexR32F>(Format::R32_SFLOAT);
            p.normalRoughness = b.take<TexNR>(Format::R10_G10_B10_A2_UNORM);
            p.penumbra = b.take<TexR16F>(Format::R16_SFLOAT);
            p.tiles = b.take<TexRG8>(Format::RG8_UNORM);
            const bool hasShadow = !first || TR;  // SIGMA_Blur.resources.hlsli:25-27
            if (hasShadow) p.shadow = b.take<typename SG::Tex>(SG::format);
            p.outPenumbra = b.take<TexR16F>(Format::R16_SFLOAT);
            p.outShadow = b.take<typename SG::Tex>(SG::format);
            if (bad(hasShadow ? 7 : 6)) return false;
            if (first && g_pendingCopy.armed) {
                if (g_pendingCopy.wide == TR) {
                    if constexpr (TR) {
                        p.copyHistory = g_pendingCopy.wideParams.history; p.copyHistoryLength = g_pendingCopy.wideParams.historyLength;
                        p.copyOutHistory = g_pendingCopy.wideParams.outHistory; p.copyOutHistoryLength = g_pendingCopy.wideParams.outHistoryLength;
                    } else {
                        p.copyHistory = g_pendingCopy.narrowParams.history; p.copyHistoryLength = g_pendingCopy.narrowParams.historyLength;
                        p.copyOutHistory = g_pendingCopy.narrowParams.outHistory; p.copyOutHistoryLength = g_pendingCopy.narrowParams.outHistoryLength;
                    }
                    p.copy = 1;
                    g_pendingCopy.armed = false;
                } else
                    sigmaFlushPendingCopy(stream);
            }
            if (first)
                launchK(sigmaBlurKernel<true, TR>, pixelGrid, block, 0, stream, cb, p, ctaY0);
            else
                launchK(sigmaBlurKernel<false, TR>, pixelGrid, block, 0, stream, cb, p, ctaY0);
            return true;
        };
        if (!(tr ? run(SigmaSignal<true>()) : run(SigmaSignal<false>()))) return (uint32_t)Result::INVALID_ARGUMENT;
    } else if (key.pass == SIGMA_TEMPORAL_STABILIZATION) {
        auto run = [&](auto sig) -> bool {
            using SG = decltype(sig);
            constexpr bool TR = std::is_same<SG, SigmaSignal<true>>::value;
            SigmaTemporalStabilizationParams<typename SG::Tex> p;
            p.viewZ = b.take<TexR32F>(Format::R32_SFLOAT);
            p.mv = b.take<TexRGBA16F>(Format::RGBA16_SFLOAT);
            p.penumbra = b.take<TexR16F>(Format::R16_SFLOAT);
            p.shadow = b.take<typename SG::Tex>(SG::format);
            p.history = b.take<typename SG::Tex>(SG::format);
            p.historyLength = b.take<TexR32U>(Format::R32_UINT);
            p.tiles = b.take<TexRG8>(Format::RG8_UNORM);
            p.outShadow = b.take<typename SG::Tex>(SG::format);
            p.outHistoryLength = b.take<TexR32U>(Format::R32_UINT);
            if (bad(9)) return false;
            launchK(sigmaTemporalStabilizationKernel<TR>, pixelGrid, block, 0, stream, cb, p, ctaY0);
            return true;
        };
        if (!(tr ? run(SigmaSignal<true>()) : run(SigmaSignal<false>()))) return (uint32_t)Result::INVALID_ARGUMENT;
    } else if (key.pass == SIGMA_SPLIT_SCREEN) {
        auto run = [&](auto sig) -> bool {
            using SG = decltype(sig);
            constexpr bool TR = std::is_same<SG, SigmaSignal<true>>::value;
            SigmaSplitScreenParams<typename SG::Tex> p = {};
            p.viewZ = b.take<TexR32F>(Format::R32_SFLOAT);
            p.penumbra = b.take<TexR16F>(Format::R16_SFLOAT);
            if (TR) p.translucency = b.take<typename SG::Tex>(SG::format);
            p.outShadow = b.take<typename SG::Tex>(SG::format);
            if (bad(TR ? 4 : 3)) return false;
            launchK(sigmaSplitScreenKernel<TR>, pixelGrid, block, 0, stream, cb, p, ctaY0);
            return true;
        };
        if (!(tr ? run(SigmaSignal<true>()) : run(SigmaSignal<false>()))) return (uint32_t)Result::INVALID_ARGUMENT;
    } else {
        err = std::string("no CUDA kernel for shader '") + id + "'";
        return (uint32_t)Result::UNSUPPORTED;
    }
    return (uint32_t)Result::SUCCESS;
}

}  // namespace nrdk
