// REBLUR_HitDistReconstruction (3x3 / 5x5): fills in the hit distance of pixels whose ray carried none (hitDist = 0) from the
// neighbours on the same surface. Reference: External/NRD/Shaders/REBLUR_HitDistReconstruction.cs.hlsl:21-167
// (RADIANCE, REBLUR_USE_DECOMPRESSED_HIT_DIST_IN_RECONSTRUCTION = 0, REBLUR_PERFORMANCE_MODE = 0; NRD_SIGNAL is a uniform run-time argument:
// a single-lobe denoiser leaves the other lobe's hit distance at 0 and its textures untouched).
// CTA = 32x8 pixels; the { normal, roughness } and { diff hitDist, spec hitDist, viewZ } of the (32 + 2B) x (8 + 2B) neighbourhood
// are staged in shared memory (B = 1 or 2), normals decoded once per texel.
#include "reblur_common.cuh"

namespace nrdk {

namespace {
constexpr int BLOCK_W = 32, BLOCK_H = 8;

template <int BORDER>
__global__ void __launch_bounds__(BLOCK_W* BLOCK_H) reblurHitDistReconstructionKernel(const __grid_constant__ ReblurConstants cb, const __grid_constant__ HitDistReconstructionParams p,
                                                                                     int signal, int occlusion, int ctaY0) {
    pdlEntry();
    // Texel access by bound format ( a uniform branch; this pass is bandwidth-bound ): the NRD_MODE = OCCLUSION permutation carries the hit distance alone
    // ( Texture2D< float >: .x of whatever is bound, REBLUR_HitDistReconstruction.cs.hlsl:151-166 ), and the RADIANCE permutation also serves
    // REBLUR_DIFFUSE_DIRECTIONAL_OCCLUSION, whose textures are RGBA16_SNORM or the application's own format
    auto ld = [&](const TexRGBA16F& t, int x, int y) -> float4 {
        if (!t.inside(x, y)) return f4(0.0f);
        if (t.fmt == (uint32_t)nrd::Format::RGBA16_SFLOAT && !occlusion) return t.fetch(x, y);
        const float4 v = anyFetch4(t, x, y);
        return occlusion ? make_float4(0.0f, 0.0f, 0.0f, v.x) : v;
    };
    auto st = [&](const TexRGBA16F& t, int x, int y, float4 v) {
        if (t.fmt == (uint32_t)nrd::Format::RGBA16_SFLOAT && !occlusion) t.store(x, y, v);
        else anyStore4(t, x, y, occlusion ? make_float4(v.w, 0.0f, 0.0f, 0.0f) : v);
    };
    const bool hasDiff = (signal & SIGNAL_DIFF) != 0, hasSpec = (signal & SIGNAL_SPEC) != 0;
    constexpr int TW = BLOCK_W + 2 * BORDER, TH = BLOCK_H + 2 * BORDER;
    __shared__ float4 sNormalRoughness[TH][TW];
    __shared__ float4 sHitDistViewZ[TH][TW];

    const int bx = blockIdx.x, by = ctaY0 + blockIdx.y;  // NRD_CTA_ORDER_DEFAULT
    const int px = bx * BLOCK_W + threadIdx.x, py = by * BLOCK_H + threadIdx.y;
    const float skyL = p.tiles.load((bx * BLOCK_W) >> 4, py >> 4), skyR = p.tiles.load((bx * BLOCK_W + 16) >> 4, py >> 4);
    if (skyL != 0.0f && skyR != 0.0f) return;
    {
        const int baseX = bx * BLOCK_W - BORDER, baseY = by * BLOCK_H - BORDER;
        for (int i = threadIdx.y * BLOCK_W + threadIdx.x; i < TW * TH; i += BLOCK_W * BLOCK_H) {
            const int tx = i % TW, ty = i / TW;
            const int gx = clampi(baseX + tx, 0, cb.rectSizeMinusOne[0]), gy = clampi(baseY + ty, 0, cb.rectSizeMinusOne[1]);
            const float viewZ = unpackViewZ(cb, p.viewZ.load(gx, gy));
            sNormalRoughness[ty][tx] = unpackNormalRoughness(p.normalRoughness.loadRaw(gx, gy));
            const bool inRange = inDenoisingRange(cb, viewZ);
            sHitDistViewZ[ty][tx] = make_float4(inRange && hasDiff ? ld(p.inDiff, gx, gy).w : 0.0f, inRange && hasSpec ? ld(p.inSpec, gx, gy).w : 0.0f, viewZ, 0.0f);
        }
    }
    __syncthreads();
    const float isSky = threadIdx.x < 16 ? skyL : skyR;
    if (isSky != 0.0f || px > cb.rectSizeMinusOne[0] || py > cb.rectSizeMinusOne[1]) return;

    const int smx = threadIdx.x + BORDER, smy = threadIdx.y + BORDER;
    const float4 center = sHitDistViewZ[smy][smx];
    if (!inDenoisingRange(cb, center.z)) return;

    const float4 nr = sNormalRoughness[smy][smx];
    const float3 N = xyz(nr);
    const float roughness = nr.w;
    const float2 rectSizeInv = make_float2(cb.rectSizeInv[0], cb.rectSizeInv[1]);
    const float2 pixelUv = make_float2((float)px + 0.5f, (float)py + 0.5f) * rectSizeInv;
    const float3 Xv = reconstructViewPosition(pixelUv, cb.frustum, center.z, cb.orthoMode);
    const float3 Nv = rotateInverse(cb.viewToWorld, N);
    const float frustumSize = frustumSizeAt(cb.minRectDimMulUnproject, cb.orthoMode, center.z);
    const float2 geomParams = geometryWeightParams(cb.planeDistSensitivity, frustumSize, Xv, Nv);
    const float2 relaxedRoughnessParams = relaxedRoughnessWeightParams(roughness * roughness);
    const float diffNormalParam = normalWeightParam(1.0f, 1.0f);
    const float specNormalParam = normalWeightParam(1.0f, 1.0f, roughness);

    float2 sum = make_float2(center.x != 0.0f ? 1000.0f : 0.0f, center.y != 0.0f ? 1000.0f : 0.0f);
    float2 acc = make_float2(center.x, center.y) * sum;
#pragma unroll
    for (int j = 0; j <= BORDER * 2; j++)
#pragma unroll
        for (int i = 0; i <= BORDER * 2; i++) {
            if (i == BORDER && j == BORDER) continue;
            const float2 o = make_float2((float)(i - BORDER), (float)(j - BORDER));
            const float4 data = sHitDistViewZ[threadIdx.y + j][threadIdx.x + i];
            const float2 uv = pixelUv + o * rectSizeInv;
            float w = isInScreenNearest(uv) ? 1.0f : 0.0f;
            w *= gaussianWeight(length(o) * 0.5f);
            // strict ( non exponential ) plane weight: no data from other surfaces
            const float3 Xvs = reconstructViewPosition(uv, cb.frustum, data.z, cb.orthoMode);
            w *= nonExponentialWeight(dot(Nv, Xvs), geomParams.x, geomParams.y);

            const float4 snr = sNormalRoughness[threadIdx.y + j][threadIdx.x + i];
            const float angle = acosApproxPositive(dot(N, xyz(snr)));
            float wd = w * exponentialWeight(angle, diffNormalParam, 0.0f);
            float ws = w * exponentialWeight(angle, specNormalParam, 0.0f);
            ws *= exponentialWeight(snr.w * snr.w, relaxedRoughnessParams.x, relaxedRoughnessParams.y);
            wd = data.x == 0.0f ? 0.0f : wd;  // ignore "no data"
            ws = data.y == 0.0f ? 0.0f : ws;
            acc.x += data.x * wd;
            acc.y += data.y * ws;
            sum.x += wd;
            sum.y += ws;
        }
    acc.x /= fmaxf(sum.x, NRD_EPS);
    acc.y /= fmaxf(sum.y, NRD_EPS);

    if (hasDiff) {
        const float4 diff = ld(p.inDiff, px, py);
        st(p.outDiff, px, py, make_float4(diff.x, diff.y, diff.z, acc.x));
    }
    if (hasSpec) {
        const float4 spec = ld(p.inSpec, px, py);
        st(p.outSpec, px, py, make_float4(spec.x, spec.y, spec.z, acc.y));
    }
}
}  // namespace

void launchReblurHitDistReconstruction(const ReblurConstants& cb, const HitDistReconstructionParams& p, int signal, bool occlusion, bool is5x5, Rows rows, cudaStream_t stream) {
    const RowGrid g = rowGrid(rows, cb.rectSizeMinusOne[1] + 1, BLOCK_H);
    if (!g.count) return;
    const dim3 grid((cb.rectSizeMinusOne[0] + BLOCK_W) / BLOCK_W, g.count), block(BLOCK_W, BLOCK_H);
    if (is5x5)
        launchK(reblurHitDistReconstructionKernel<2>, grid, block, 0, stream, cb, p, signal, occlusion ? 1 : 0, g.ctaY0);
    else
        launchK(reblurHitDistReconstructionKernel<1>, grid, block, 0, stream, cb, p, signal, occlusion ? 1 : 0, g.ctaY0);
}

}  // namespace nrdk
