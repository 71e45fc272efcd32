// TEST INFRASTRUCTURE — CPU oracle (see hlsl_like.h). NOT part of the product.
// PINNED: bit-identical, dispatch by dispatch, to the reference's own shaders compiled as C++
// (oracle/_ref/libnrd_refshaders.so, tests/test_oracle_vs_reference_shaders.py).
//
// SIGMA_SHADOW / SIGMA_SHADOW_TRANSLUCENCY (TRANSLUCENCY = 0 / 1: SIGMA_TYPE = float / float4, SIGMA_Config.hlsli:37-42) restated from /root/reference/External/NRD/Shaders:
//   SIGMA_ClassifyTiles.cs.hlsl:24-91, SIGMA_SmoothTiles.cs.hlsl:21-58, SIGMA_Copy.cs.hlsl:19-32,
//   SIGMA_Blur.cs.hlsl:21-286 (FIRST_PASS = 1 / 0), SIGMA_TemporalStabilization.cs.hlsl:21-236,
//   SIGMA_SplitScreen.cs.hlsl:21-45, helpers SIGMA_Common.hlsli:13-95, switches SIGMA_Config.hlsli:11-42.
// Shared memory tiles of the shaders hold f( clamp( pos, 0, rectSizeMinusOne ) ); the restatement reads the
// clamped texel directly. Scalar loops, OpenMP over rows.
#include <cmath>
#include <string>

#include "reblur_shared.h"

namespace orc {

namespace {

struct SigmaCB {  // SIGMA_Config.hlsli:44-78 (HLSL cbuffer packing), 528 bytes
    float4x4 gWorldToView, gViewToClip, gWorldToClipPrev, gWorldToViewPrev;
    float4 gRotator, gRotatorPost, gViewVectorWorld, gLightDirectionView, gFrustum, gFrustumPrev, gCameraDelta, gMvScale;
    float2 gResourceSizeInv, gResourceSizeInvPrev, gRectSize, gRectSizeInv, gRectSizePrev, gResolutionScale, gRectOffset;
    uint2 gPrintfAt, gRectOrigin;
    int2 gRectSizeMinusOne, gTilesSizeMinusOne;
    float gOrthoMode, gUnproject, gDenoisingRange, gPlaneDistSensitivity, gStabilizationStrength, gDebug, gSplitScreen, gViewZScale, gMinRectDimMulUnproject;
    uint32_t gFrameIndex, gIsRectChanged;
    uint32_t _pad[3];
};
static_assert(sizeof(SigmaCB) == 528, "SIGMA cbuffer is 528 bytes");

const float SIGMA_MAX_PIXEL_RADIUS = 32.0f;  // SIGMA_Config.hlsli:33
const float SIGMA_TS_SIGMA_SCALE = 3.0f;     // :34
const float SIGMA_MAX_ACCUM_FRAME_NUM = 7;   // :35
const int BLUR_BORDER = 2;                   // SIGMA_5X5_BLUR_RADIUS_ESTIMATION_KERNEL = 1
const int TS_BORDER = 2;                     // SIGMA_5X5_TEMPORAL_KERNEL = 1

inline float PackShadow(float s) { return Math::Sqrt01(s); }     // SIGMA_Common.hlsli:13
inline bool IsLit(float p) { return p >= NRD_FP16_MAX; }         // :14
inline float UnpackShadow(float s) { return s * s; }             // NRD.hlsli:1010
inline float4 PackShadow(float4 s) { return float4(Math::Sqrt01(s.x), Math::Sqrt01(s.y), Math::Sqrt01(s.z), Math::Sqrt01(s.w)); }
inline float4 UnpackShadow(float4 s) { return s * s; }
inline float StdDevS(float m1, float m2) { return GetStdDev(m1, m2); }
inline float4 StdDevS(float4 m1, float4 m2) { return float4(GetStdDev(m1.x, m2.x), GetStdDev(m1.y, m2.y), GetStdDev(m1.z, m2.z), GetStdDev(m1.w, m2.w)); }
inline float ClampS(float x, float lo, float hi) { return clamp(x, lo, hi); }
inline float4 ClampS(float4 x, float4 lo, float4 hi) { return min(max(x, lo), hi); }

// SIGMA_TYPE (SIGMA_Config.hlsli:37-42): float, or float4 = { shadow, translucency.rgb }
template <bool TRANSLUCENCY> struct Sig;
template <> struct Sig<false> {
    using T = float;
    static T get(float4 texel) { return texel.x; }
    static float4 texel(T s) { return float4(s, 0, 0, 0); }
    static T splat(float v) { return v; }
    static float x(T s) { return s; }
};
template <> struct Sig<true> {
    using T = float4;
    static T get(float4 texel) { return texel; }
    static float4 texel(T s) { return s; }
    static T splat(float v) { return float4(v); }
    static float x(T s) { return s.x; }
};

// SIGMA_Common.hlsli:21-34. min / clamp are IEEE minNum / maxNum like the HLSL intrinsics: 0 / 0 (texels outside the
// resource read viewZ = 0) collapses to the lower bound instead of propagating the NaN
inline float GetKernelRadiusInPixels(float hitDist, float unprojectZ, float scale = 1.0f) {
    float unclampedRadius = hitDist / unprojectZ;
    unclampedRadius *= scale;
    float minRadius = std::fmin(unclampedRadius, 2.0f);
    return std::fmin(std::fmax(unclampedRadius, minRadius), SIGMA_MAX_PIXEL_RADIUS);
}
inline float AreBothLitOrUnlit(float p1, float p2) { return float((p1 == 0.0f) == (p2 == 0.0f)); }  // :36-42

struct Ctx {
    const SigmaCB& cb;
    explicit Ctx(const SigmaCB& c) : cb(c) {}
    float UnpackViewZ(float z) const { return std::fabs(z * cb.gViewZScale); }
    bool IsInDenoisingRange(float z) const { return z < cb.gDenoisingRange; }
    float ApplyGeometryWeightLast(float w, float z, float NoX, float2 p) const {  // Common.hlsli:567
        w *= ComputeWeight(NoX, p.x, p.y);
        return !IsInDenoisingRange(z) ? 0.0f : w;
    }
    float3 GetViewVector(float3 X, bool isViewSpace) const {  // SIGMA_Common.hlsli:16-19
        return cb.gOrthoMode == 0.0f ? normalize(-X) : (isViewSpace ? float3(0, 0, -1) : cb.gViewVectorWorld.xyz());
    }
};

// SIGMA_Common.hlsli:46-75, returns ( yw.z, xw.z )
float2 FilterBicubic(float2 size, float2 uv, float4& uv_10_00, float4& uv_11_01) {
    const float k = 1.0f / 6.0f;
    float4 dxdy = float4(-1.0f / size.x, -0.0f / size.y, -0.0f / size.x, -1.0f / size.y);
    float2 f = frac(uv * size - 0.5f);
    float2 f2 = f * f;
    float2 f3 = f2 * f;
    auto axis = [&](float fa, float fa2, float fa3) {
        float4 phi;
        phi.x = k * (-1.0f * fa3 + 3.0f * fa2 + -3.0f * fa + 1.0f);
        phi.y = k * (3.0f * fa3 + -6.0f * fa2 + 0.0f * fa + 4.0f);
        phi.z = k * (-3.0f * fa3 + 3.0f * fa2 + 3.0f * fa + 1.0f);
        phi.w = k * (1.0f * fa3 + 0.0f * fa2 + 0.0f * fa + 0.0f);
        float3 r;
        r.x = 1.0f + 1.0f * fa + -1.0f * phi.y / (phi.x + phi.y);
        r.y = 1.0f + -1.0f * fa + 1.0f * phi.w / (phi.z + phi.w);
        r.z = phi.x + phi.y;
        return r;
    };
    float3 xw = axis(f.x, f2.x, f3.x);
    float3 yw = axis(f.y, f2.y, f3.y);
    uv_10_00 = float4(uv.x + 1.0f * xw.x * dxdy.x, uv.y + 1.0f * xw.x * dxdy.y, uv.x + -1.0f * xw.y * dxdy.x, uv.y + -1.0f * xw.y * dxdy.y);
    uv_11_01 = uv_10_00 + float4(yw.x * dxdy.z, yw.x * dxdy.w, yw.x * dxdy.z, yw.x * dxdy.w);
    uv_10_00 = uv_10_00 - float4(yw.y * dxdy.z, yw.y * dxdy.w, yw.y * dxdy.z, yw.y * dxdy.w);
    return float2(yw.z, xw.z);
}

// SIGMA_Common.hlsli:77-95 on the RG8 smoothed-tiles texture
float2 TextureCubic(const Tex& tex, float2 uv) {
    float2 size = float2((float)tex.w, (float)tex.h);
    float4 uv_10_00, uv_11_01;
    float2 t = FilterBicubic(size, uv, uv_10_00, uv_11_01);
    float4 c00 = tex.sampleLinear(float2(uv_10_00.z, uv_10_00.w));
    float4 c10 = tex.sampleLinear(float2(uv_10_00.x, uv_10_00.y));
    float4 c01 = tex.sampleLinear(float2(uv_11_01.z, uv_11_01.w));
    float4 c11 = tex.sampleLinear(float2(uv_11_01.x, uv_11_01.y));
    c00 = lerp(c00, c01, t.x);
    c10 = lerp(c10, c11, t.x);
    float4 r = lerp(c00, c10, t.y);
    return float2(r.x, r.y);
}

inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// ---------------------------------------------------------------------------------------------------------------
// SIGMA_ClassifyTiles.cs.hlsl:24-91 — one 16x16 tile per group
template <bool TRANSLUCENCY>
void classifyTiles(const SigmaCB& cb, const Tex& gIn_ViewZ, const Tex& gIn_Penumbra, const Tex* gIn_Shadow_Translucency, Tex& gOut_Tiles, int gridW, int gridH) {
    Ctx c(cb);
#pragma omp parallel for schedule(dynamic, 1)
    for (int ty = 0; ty < gridH; ty++)
        for (int tx = 0; tx < gridW; tx++) {
            uint32_t nLit = 0, nUmbra = 0, nInf = 0;
            float maxRadius = 0.0f;
            for (int j = 0; j < 16; j++)
                for (int i = 0; i < 16; i++) {
                    int px = tx * 16 + i, py = ty * 16 + j;
                    float h = gIn_Penumbra.load(px, py).x;
                    float viewZ = c.UnpackViewZ(gIn_ViewZ.load(px, py).x);
                    bool isInf = !c.IsInDenoisingRange(viewZ);
                    bool isShadow = h == 0.0f;
                    bool isLit = IsLit(h);
                    bool isOpaque = true;
                    if (TRANSLUCENCY) {
                        float4 st = gIn_Shadow_Translucency->load(px, py);
                        isOpaque = Color::Luminance(float3(st.y, st.z, st.w)) < 0.003f;
                    }
                    nLit += (isLit || isInf || isShadow) ? 1 : 0;
                    nUmbra += ((!isLit && isOpaque) || isInf || isShadow) ? 1 : 0;
                    nInf += isInf ? 1 : 0;
                    float hitDist = (isLit || isInf) ? 0.0f : h;
                    float pixelSize = PixelRadiusToWorld(cb.gUnproject, cb.gOrthoMode, 1.0f, viewZ);
                    float pixelRadius = GetKernelRadiusInPixels(hitDist, pixelSize);
                    maxRadius = std::fmax(pixelRadius, maxRadius);
                }
            // 9-bit fields of s_Mask hold the counts; "== 256" = every pixel of the tile
            bool isLit = (nLit & 511u) == 256u, isUmbra = (nUmbra & 511u) == 256u, isInf = (nInf & 511u) == 256u;
            float4 result;
            result.x = (isLit || isUmbra) ? 0.0f : 1.0f;
            result.y = saturate(maxRadius / 16.0f);
            result.z = isInf ? 1.0f : 0.0f;
            result.w = 0.0f;
            gOut_Tiles.store(tx, ty, result);
        }
}

// SIGMA_SmoothTiles.cs.hlsl:21-58 — one thread per tile texel, 3x3 neighbourhood clamped to the tile grid
void smoothTiles(const SigmaCB& cb, const Tex& gIn_Tiles, Tex& gOut_Tiles, int gridW, int gridH) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int y = 0; y < gridH * 16; y++)
        for (int x = 0; x < gridW * 16; x++) {
            float4 center = gIn_Tiles.load(x, y);
            float blurry = 0.0f, sum = 0.0f;
            float k = 1.01f / (center.y + 0.01f);
            for (int j = 0; j <= 2; j++)
                for (int i = 0; i <= 2; i++) {
                    float d = length(float2((float)i, (float)j) - 1.0f);
                    float w = std::exp2(-k * d * d);
                    int gx = clampi(x + i - 1, 0, cb.gTilesSizeMinusOne.x), gy = clampi(y + j - 1, 0, cb.gTilesSizeMinusOne.y);
                    blurry += gIn_Tiles.load(gx, gy).x * w;
                    sum += w;
                }
            blurry /= sum;
            gOut_Tiles.store(x, y, float4(center.z, blurry, 0.0f, 0.0f));
        }
}

// SIGMA_Copy.cs.hlsl:19-32
void copy(const SigmaCB& cb, const Tex& gIn_Tiles, const Tex& gIn_History, const Tex& gIn_HistoryLength, Tex& gOut_History, Tex& gOut_HistoryLength, int gridW,
          int gridH) {
#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < gridH * 16; y++)
        for (int x = 0; x < gridW * 8; x++) {
            float isSky = gIn_Tiles.load(x >> 4, y >> 4).x;
            if (isSky != 0.0f && !cb.gIsRectChanged) continue;
            gOut_History.store(x, y, gIn_History.load(x, y));
            gOut_HistoryLength.storeUint(x, y, gIn_HistoryLength.loadUint(x, y));
        }
}

// SIGMA_Blur.cs.hlsl:21-286
template <bool TRANSLUCENCY>
void blur(const SigmaCB& cb, bool firstPass, const Tex& gIn_ViewZ, const Tex& gIn_Normal_Roughness, const Tex& gIn_Penumbra, const Tex& gIn_Tiles,
          const Tex* gIn_Shadow_Translucency, Tex& gOut_Penumbra, Tex& gOut_Shadow_Translucency, int gridW, int gridH) {
    using SG = Sig<TRANSLUCENCY>;
    using S = typename SG::T;
    Ctx c(cb);
    // Preload( ) of the shader: { penumbra, viewZ }, shadow at the rect-clamped position
    auto preloadPV = [&](int x, int y) {
        int gx = clampi(x, 0, cb.gRectSizeMinusOne.x), gy = clampi(y, 0, cb.gRectSizeMinusOne.y);
        return float2(gIn_Penumbra.load(gx, gy).x, c.UnpackViewZ(gIn_ViewZ.load(gx, gy).x));
    };
    auto preloadS = [&](int x, int y, float penumbra) -> S {
        int gx = clampi(x, 0, cb.gRectSizeMinusOne.x), gy = clampi(y, 0, cb.gRectSizeMinusOne.y);
        if (firstPass && !TRANSLUCENCY) return SG::splat(float(IsLit(penumbra)));
        S s = SG::get(gIn_Shadow_Translucency->load(gx, gy));
        return firstPass ? s : UnpackShadow(s);
    };
#pragma omp parallel for schedule(dynamic, 4)
    for (int py = 0; py < gridH * 16; py++)
        for (int px = 0; px < gridW * 8; px++) {
            float isSky = gIn_Tiles.load(px >> 4, py >> 4).x;
            if (isSky != 0.0f || px > cb.gRectSizeMinusOne.x || py > cb.gRectSizeMinusOne.y) continue;

            float2 centerData = preloadPV(px, py);
            float centerPenumbra = centerData.x;
            float viewZ = centerData.y;
            if (!c.IsInDenoisingRange(viewZ)) continue;

            float2 pixelUv = float2(px + 0.5f, py + 0.5f) * cb.gRectSizeInv;
            float tileValue = TextureCubic(gIn_Tiles, pixelUv * cb.gResolutionScale).y;

            if (tileValue == 0.0f || centerPenumbra == 0.0f) {
                if (firstPass || cb.gStabilizationStrength != 0.0f) gOut_Penumbra.store(px, py, float4(centerPenumbra, 0, 0, 0));
                gOut_Shadow_Translucency.store(px, py, SG::texel(PackShadow(preloadS(px, py, centerPenumbra))));
                continue;
            }

            float3 Xv = Geometry::ReconstructViewPosition(pixelUv, cb.gFrustum, viewZ, cb.gOrthoMode);
            float4 normalAndRoughness = NRD_FrontEnd_UnpackNormalAndRoughness(gIn_Normal_Roughness.load(px, py));
            float3 N = normalAndRoughness.xyz();
            float3 Nv = Geometry::RotateVector(cb.gWorldToView, N);

            float pixelSize = PixelRadiusToWorld(cb.gUnproject, cb.gOrthoMode, 1.0f, viewZ);
            float frustumSize = GetFrustumSize(cb.gMinRectDimMulUnproject, cb.gOrthoMode, viewZ);
            float3 Vv = c.GetViewVector(Xv, true);
            float NoV = std::fabs(dot(Nv, Vv));
            float2 geometryWeightParams = GetGeometryWeightParams(cb.gPlaneDistSensitivity, frustumSize, Xv, Nv);

            // Estimate penumbra size and filter shadow ( dense )
            float2 sum = float2(0.0f);
            float penumbra = 0.0f;
            S result = SG::splat(0.0f), centerTap = SG::splat(0.0f);
            for (int j = 0; j <= BLUR_BORDER * 2; j++)
                for (int i = 0; i <= BLUR_BORDER * 2; i++) {
                    int sx = px + i - BLUR_BORDER, sy = py + j - BLUR_BORDER;
                    float2 data = preloadPV(sx, sy);
                    float penum = data.x, zs = data.y;
                    S s = preloadS(sx, sy, penum);
                    float w = 1.0f;
                    if (i == BLUR_BORDER && j == BLUR_BORDER)
                        centerTap = s;
                    else {
                        float2 o = float2(float(i - BLUR_BORDER), float(j - BLUR_BORDER));
                        float2 uv = pixelUv + o * cb.gRectSizeInv;
                        float3 Xvs = Geometry::ReconstructViewPosition(uv, cb.gFrustum, zs, cb.gOrthoMode);
                        float NoX = dot(Nv, Xvs);
                        w *= AreBothLitOrUnlit(centerPenumbra, penum);
                        w *= GetGaussianWeight(length(o / float(BLUR_BORDER)));
                        w = c.ApplyGeometryWeightLast(w, zs, NoX, geometryWeightParams);
                    }
                    result = result + (w == 0.0f ? SG::splat(0.0f) : s * w);
                    sum.x += w;
                    w *= pixelSize / (pixelSize + penum);
                    w *= float(!IsLit(penum));
                    penumbra += w == 0.0f ? 0.0f : penum * w;
                    sum.y += w;
                }
            result = result / sum.x;
            sum.x = 1.0f;
            penumbra /= max(sum.y, NRD_EPS);
            sum.y = float(sum.y != 0.0f);

            // Avoid blurry result if penumbra size < NRD_BORDER
            float penumbraInPixels = penumbra / pixelSize;
            float f = Math::SmoothStep(0.0f, float(BLUR_BORDER), penumbraInPixels);
            result = lerp(centerTap, result, f);

            // SIGMA_USE_SPARSE_BLUR = 1
            f = lerp(4.0f, 1.0f, f);
            result = result * f;
            penumbra *= f;
            sum = sum * f;

            float blurRadius = GetKernelRadiusInPixels(penumbra, pixelSize, tileValue);
            float4 rotator = firstPass ? cb.gRotator : cb.gRotatorPost;  // SIGMA_ROTATOR_MODE = NRD_FRAME

            // SIGMA_USE_SCREEN_SPACE_SAMPLING = 1
            float2 skew = lerp(float2(1.0f) - abs(float2(Nv.x, Nv.y)), float2(1.0f), NoV);
            skew /= max(skew.x, skew.y);
            skew *= cb.gRectSizeInv * blurRadius;
            float4 scaledRotator = Geometry::ScaleRotator(rotator, skew);

            float invEstimatedPenumbra = 1.0f / max(penumbra, NRD_EPS);
            for (int n = 0; n < 8; n++) {
                float3 offset = g_Special8[n];
                float2 uv = pixelUv + Geometry::RotateVector(scaledRotator, float2(offset.x, offset.y));
                uv = (floor(uv * cb.gRectSize) + 0.5f) * cb.gRectSizeInv;  // snap to the pixel center
                // ClampUvToViewport, NRD_SUPPORTS_VIEWPORT_OFFSET = 0 (Common.hlsli:242)
                float2 uvScaled = min(uv * cb.gResolutionScale, cb.gResolutionScale - 0.5f * cb.gResourceSizeInv);

                float penum = gIn_Penumbra.sampleNearest(uvScaled).x;
                float zs = c.UnpackViewZ(gIn_ViewZ.sampleNearest(uvScaled).x);
                float3 Xvs = Geometry::ReconstructViewPosition(uv, cb.gFrustum, zs, cb.gOrthoMode);
                S s;
                if (firstPass && !TRANSLUCENCY)
                    s = SG::splat(float(IsLit(penum)));
                else {
                    s = SG::get(gIn_Shadow_Translucency->sampleNearest(uvScaled));
                    if (!firstPass) s = UnpackShadow(s);
                }

                float NoX = dot(Nv, Xvs);
                float w = IsInScreenNearest(uv);
                w *= AreBothLitOrUnlit(centerPenumbra, penum);
                w *= GetGaussianWeight(offset.z);
                w *= saturate(penum * invEstimatedPenumbra);  // avoid umbra leaking inside wide penumbra
                w = c.ApplyGeometryWeightLast(w, zs, NoX, geometryWeightParams);

                result = result + (w == 0.0f ? SG::splat(0.0f) : s * w);
                sum.x += w;
                w *= pixelSize / (pixelSize + penum);
                w *= float(!IsLit(penum));
                penumbra += w == 0.0f ? 0.0f : penum * w;
                sum.y += w;
            }

            result = result / sum.x;
            penumbra = sum.y == 0.0f ? centerPenumbra : penumbra / sum.y;

            if (firstPass || cb.gStabilizationStrength != 0.0f) gOut_Penumbra.store(px, py, float4(penumbra, 0, 0, 0));
            gOut_Shadow_Translucency.store(px, py, SG::texel(PackShadow(result)));
        }
}

inline uint32_t PackViewZAndHistoryLength(float viewZ, float historyLength) {  // SIGMA_TemporalStabilization.cs.hlsl:36-42
    uint32_t p = asuint(viewZ) & ~7u;
    uint32_t n = (uint32_t)(historyLength + 0.5f);
    p |= n < 7u ? n : 7u;
    return p;
}

// SIGMA_TemporalStabilization.cs.hlsl:53-236
template <bool TRANSLUCENCY>
void temporalStabilization(const SigmaCB& cb, const Tex& gIn_ViewZ, const Tex& gIn_Mv, const Tex& gIn_Penumbra, const Tex& gIn_Shadow_Translucency,
                           const Tex& gIn_History, const Tex& gIn_HistoryLength, const Tex& gIn_Tiles, Tex& gOut_Shadow_Translucency, Tex& gOut_HistoryLength,
                           int gridW, int gridH) {
    using SG = Sig<TRANSLUCENCY>;
    using S = typename SG::T;
    Ctx c(cb);
    auto preloadS = [&](int x, int y) -> S {
        int gx = clampi(x, 0, cb.gRectSizeMinusOne.x), gy = clampi(y, 0, cb.gRectSizeMinusOne.y);
        return UnpackShadow(SG::get(gIn_Shadow_Translucency.load(gx, gy)));
    };
    auto preloadP = [&](int x, int y) {
        int gx = clampi(x, 0, cb.gRectSizeMinusOne.x), gy = clampi(y, 0, cb.gRectSizeMinusOne.y);
        return gIn_Penumbra.load(gx, gy).x;
    };
#pragma omp parallel for schedule(dynamic, 4)
    for (int py = 0; py < gridH * 16; py++)
        for (int px = 0; px < gridW * 8; px++) {
            float isSky = gIn_Tiles.load(px >> 4, py >> 4).x;
            float viewZ = c.UnpackViewZ(gIn_ViewZ.load(px, py).x);
            if (isSky != 0.0f || px > cb.gRectSizeMinusOne.x || py > cb.gRectSizeMinusOne.y || !c.IsInDenoisingRange(viewZ)) continue;
            float centerPenumbra = preloadP(px, py);

            float2 pixelUv = float2(px + 0.5f, py + 0.5f) * cb.gRectSizeInv;
            float tileValue = TextureCubic(gIn_Tiles, pixelUv * cb.gResolutionScale).y;
            bool isHardShadow = tileValue == 0.0f || centerPenumbra == 0.0f;  // NRD_USE_TILE_CHECK = SIGMA_USE_EARLY_OUT_IN_TS = 1
            if (isHardShadow) {
                gOut_Shadow_Translucency.store(px, py, SG::texel(PackShadow(preloadS(px, py))));
                gOut_HistoryLength.storeUint(px, py, PackViewZAndHistoryLength(viewZ, SIGMA_MAX_ACCUM_FRAME_NUM));
                continue;
            }

            // Local variance
            float sum = 0.0f;
            S m1 = SG::splat(0.0f), m2 = SG::splat(0.0f), input = SG::splat(0.0f);
            for (int j = 0; j <= TS_BORDER * 2; j++)
                for (int i = 0; i <= TS_BORDER * 2; i++) {
                    int sx = px + i - TS_BORDER, sy = py + j - TS_BORDER;
                    S s = preloadS(sx, sy);
                    float w = 1.0f;
                    if (i == TS_BORDER && j == TS_BORDER)
                        input = s;
                    else {
                        float penum = preloadP(sx, sy);
                        w = AreBothLitOrUnlit(centerPenumbra, penum);
                        w *= GetGaussianWeight(length(float2(float(i - TS_BORDER), float(j - TS_BORDER)) / float(TS_BORDER)));
                    }
                    m1 = m1 + s * w;
                    m2 = m2 + s * s * w;
                    sum += w;
                }
            m1 = m1 / sum;
            m2 = m2 / sum;
            S sigma = StdDevS(m1, m2);

            // Current and previous positions
            float3 Xv = Geometry::ReconstructViewPosition(pixelUv, cb.gFrustum, viewZ, cb.gOrthoMode);
            float3 X = Geometry::RotateVectorInverse(cb.gWorldToView, Xv);
            float4 mvRaw = gIn_Mv.load(px, py);
            float3 mv = float3(mvRaw.x, mvRaw.y, mvRaw.z) * cb.gMvScale.xyz();
            float3 Xprev = X;
            float2 smbPixelUv = pixelUv + float2(mv.x, mv.y);
            if (cb.gMvScale.w == 0.0f) {
                if (cb.gMvScale.z == 0.0f) mv.z = Geometry::AffineTransform(cb.gWorldToViewPrev, X).z - viewZ;
                float viewZprev = viewZ + mv.z;
                float3 Xvprevlocal = Geometry::ReconstructViewPosition(smbPixelUv, cb.gFrustumPrev, viewZprev, cb.gOrthoMode);
                Xprev = Geometry::RotateVectorInverse(cb.gWorldToViewPrev, Xvprevlocal) + cb.gCameraDelta.xyz();
            } else {
                Xprev = Xprev + mv;
                smbPixelUv = Geometry::GetScreenUv(cb.gWorldToClipPrev, Xprev);
            }

            // History length
            Filtering::Bilinear smbBilinearFilter = Filtering::GetBilinearFilter(smbPixelUv, cb.gRectSizePrev);
            // gather uv = ( origin + 1 ) * invSize -> footprint top-left texel = origin (clamp addressing)
            int ox = (int)smbBilinearFilter.origin.x, oy = (int)smbBilinearFilter.origin.y;
            uint32_t prevData[4] = {gIn_HistoryLength.fetchUintClamped(ox, oy), gIn_HistoryLength.fetchUintClamped(ox + 1, oy),
                                    gIn_HistoryLength.fetchUintClamped(ox, oy + 1), gIn_HistoryLength.fetchUintClamped(ox + 1, oy + 1)};
            float4 prevViewZ = float4(asfloat(prevData[0] & ~7u), asfloat(prevData[1] & ~7u), asfloat(prevData[2] & ~7u), asfloat(prevData[3] & ~7u));
            float4 prevHistoryLength = float4(float(prevData[0] & 7u), float(prevData[1] & 7u), float(prevData[2] & 7u), float(prevData[3] & 7u));

            float frustumSize = GetFrustumSize(cb.gMinRectDimMulUnproject, cb.gOrthoMode, viewZ);
            float4 disocclusionThreshold = float4(GetDisocclusionThreshold(NRD_DISOCCLUSION_THRESHOLD, frustumSize, 1.0f));
            disocclusionThreshold = disocclusionThreshold * IsInScreenBilinear(smbBilinearFilter.origin, cb.gRectSizePrev);
            disocclusionThreshold -= NRD_EPS;

            float3 Xvprev = Geometry::AffineTransform(cb.gWorldToViewPrev, Xprev);
            float4 smbPlaneDist = abs(prevViewZ - float4(Xvprev.z));
            float4 smbOcclusion = step(smbPlaneDist, disocclusionThreshold);
            float4 smbOcclusionWeights = Filtering::GetBilinearCustomWeights(smbBilinearFilter, smbOcclusion);
            float historyLength = Filtering::ApplyBilinearCustomWeights(prevHistoryLength.x, prevHistoryLength.y, prevHistoryLength.z, prevHistoryLength.w,
                                                                        smbOcclusionWeights);

            // Sample history
            bool isCatRomAllowed = dot(smbOcclusionWeights, float4(1.0f)) > 3.5f;
            HistoryFilter hf(saturate(smbPixelUv) * cb.gRectSizePrev, cb.gResourceSizeInvPrev, smbOcclusionWeights, isCatRomAllowed);
            S history = SG::get(hf.color(gIn_History));
            history = saturate(history);
            history = UnpackShadow(history);

            // Clamp history
            sigma = sigma * lerp(SIGMA_TS_SIGMA_SCALE, 1.0f, 1.0f / (1.0f + historyLength));
            S inputMin = m1 - sigma, inputMax = m1 + sigma;
            S historyClamped = ClampS(history, inputMin, inputMax);

            // Antilag ( SIGMA_ADJUST_HISTORY_LENGTH_BY_ANTILAG = 1 )
            float antilag = std::fabs(SG::x(historyClamped) - SG::x(history));
            antilag = Math::Sqrt01(antilag);
            antilag = saturate(1.0f - antilag);
            historyLength *= antilag;

            float historyWeight = historyLength / (1.0f + historyLength);
            float streetMagic = 0.6f * historyWeight * antilag;
            historyClamped = lerp(historyClamped, history, streetMagic);

            S result = lerp(input, historyClamped, min(cb.gStabilizationStrength, historyWeight));
            historyLength = min(historyLength + 1.0f, SIGMA_MAX_ACCUM_FRAME_NUM);

            gOut_Shadow_Translucency.store(px, py, SG::texel(PackShadow(result)));
            gOut_HistoryLength.storeUint(px, py, PackViewZAndHistoryLength(viewZ, historyLength));
        }
}

// SIGMA_SplitScreen.cs.hlsl:21-45
template <bool TRANSLUCENCY>
void splitScreen(const SigmaCB& cb, const Tex& gIn_ViewZ, const Tex& gIn_Penumbra, const Tex* gIn_Shadow_Translucency, Tex& gOut_Shadow_Translucency, int gridW, int gridH) {
    using SG = Sig<TRANSLUCENCY>;
    Ctx c(cb);
    for (int py = 0; py < gridH * 16; py++)
        for (int px = 0; px < gridW * 8; px++) {
            float2 pixelUv = float2(px + 0.5f, py + 0.5f) * cb.gRectSizeInv;
            if (pixelUv.x > cb.gSplitScreen || px > cb.gRectSizeMinusOne.x || py > cb.gRectSizeMinusOne.y) continue;
            float viewZ = c.UnpackViewZ(gIn_ViewZ.load(px, py).x);
            typename SG::T s = TRANSLUCENCY ? SG::get(gIn_Shadow_Translucency->load(px, py)) : SG::splat(float(IsLit(gIn_Penumbra.load(px, py).x)));
            gOut_Shadow_Translucency.store(px, py, SG::texel(s * float(c.IsInDenoisingRange(viewZ))));
        }
}

}  // namespace

// returns 0 on success, 1 unknown shader, 2 bad arguments (same contract as nrd_oracle_dispatch)
int sigmaDispatch(const std::string& id, const void* constants, uint32_t cbSize, Tex* t, uint32_t n, int gridW, int gridH) {
    if (cbSize != sizeof(SigmaCB)) return 2;
    const SigmaCB& cb = *(const SigmaCB*)constants;
    if (id == "SIGMA_ClassifyTiles.cs.hlsl|TRANSLUCENCY=0") {
        if (n != 3) return 2;
        classifyTiles<false>(cb, t[0], t[1], nullptr, t[2], gridW, gridH);
        return 0;
    }
    if (id == "SIGMA_ClassifyTiles.cs.hlsl|TRANSLUCENCY=1") {
        if (n != 4) return 2;
        classifyTiles<true>(cb, t[0], t[1], &t[2], t[3], gridW, gridH);
        return 0;
    }
    if (id == "SIGMA_SmoothTiles.cs.hlsl") {
        if (n != 2) return 2;
        smoothTiles(cb, t[0], t[1], gridW, gridH);
        return 0;
    }
    if (id == "SIGMA_Copy.cs.hlsl") {
        if (n != 5) return 2;
        copy(cb, t[0], t[1], t[2], t[3], t[4], gridW, gridH);
        return 0;
    }
    if (id == "SIGMA_Blur.cs.hlsl|TRANSLUCENCY=0|FIRST_PASS=1") {
        if (n != 6) return 2;
        blur<false>(cb, true, t[0], t[1], t[2], t[3], nullptr, t[4], t[5], gridW, gridH);
        return 0;
    }
    if (id == "SIGMA_Blur.cs.hlsl|TRANSLUCENCY=0|FIRST_PASS=0") {
        if (n != 7) return 2;
        blur<false>(cb, false, t[0], t[1], t[2], t[3], &t[4], t[5], t[6], gridW, gridH);
        return 0;
    }
    if (id == "SIGMA_Blur.cs.hlsl|TRANSLUCENCY=1|FIRST_PASS=1" || id == "SIGMA_Blur.cs.hlsl|TRANSLUCENCY=1|FIRST_PASS=0") {
        if (n != 7) return 2;
        blur<true>(cb, id.back() == '1', t[0], t[1], t[2], t[3], &t[4], t[5], t[6], gridW, gridH);
        return 0;
    }
    if (id == "SIGMA_TemporalStabilization.cs.hlsl|TRANSLUCENCY=0" || id == "SIGMA_TemporalStabilization.cs.hlsl|TRANSLUCENCY=1") {
        if (n != 9) return 2;
        if (id.back() == '1')
            temporalStabilization<true>(cb, t[0], t[1], t[2], t[3], t[4], t[5], t[6], t[7], t[8], gridW, gridH);
        else
            temporalStabilization<false>(cb, t[0], t[1], t[2], t[3], t[4], t[5], t[6], t[7], t[8], gridW, gridH);
        return 0;
    }
    if (id == "SIGMA_SplitScreen.cs.hlsl|TRANSLUCENCY=0") {
        if (n != 3) return 2;
        splitScreen<false>(cb, t[0], t[1], nullptr, t[2], gridW, gridH);
        return 0;
    }
    if (id == "SIGMA_SplitScreen.cs.hlsl|TRANSLUCENCY=1") {
        if (n != 4) return 2;
        splitScreen<true>(cb, t[0], t[1], &t[2], t[3], gridW, gridH);
        return 0;
    }
    return 1;
}

}  // namespace orc
