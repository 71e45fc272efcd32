// TEST INFRASTRUCTURE — CPU oracle (see hlsl_like.h).
// Restatement of the shared shader libraries the REBLUR / SIGMA passes include:
//   MathLib   External/NRIFramework/External/MathLib/ml.hlsli   (cited as ml:LINE)
//   NRD.hlsli External/NRD/Shaders/NRD.hlsli                    (cited as nrd:LINE)
//   Common    External/NRD/Shaders/Common.hlsli                 (cited as common:LINE)
// Configuration in force (NRD/CMakeLists.txt:14-30): NRD_NORMAL_ENCODING=2 (R10G10B10A2), NRD_ROUGHNESS_ENCODING=1
// (linear), viewport offset off, checkerboard/confidence/threshold-mix/anti-firefly/quad intrinsics supported.
#pragma once
#include "texture.h"

namespace orc {

static const float NRD_EPS = 1e-6f;             // nrd:357
static const float NRD_INF = 1e6f;              // nrd:358
static const float NRD_FP16_MAX = 65504.0f;     // nrd:355
static const float NRD_NORMAL_ENCODING_ERROR = 0.75f / 255.0f;  // common:75 (R10G10B10A2)
static const float NRD_ROUGHNESS_SENSITIVITY = 0.01f;           // common:66
static const float NRD_MAX_PERCENT_OF_LOBE_VOLUME = 0.75f;      // common:69
static const float NRD_CATROM_SHARPNESS = 0.5f;                 // common:63
static const float ML_SMALL_EPS = 1e-15f, ML_EPS = 1e-6f;       // ml:31-32

// ------------------------------------------------------------------------------------------------------------
// Math (ml:58-400)
// ------------------------------------------------------------------------------------------------------------
namespace Math {
inline float Pi(float x) { return x * 3.14159265358979323846f; }
inline float DegToRad(float x) { return x * (3.14159265358979323846f / 180.0f); }
inline float LinearStep(float a, float b, float x) { return saturate((x - a) / (b - a)); }                    // ml:106
inline float SmoothStep01(float x) { x = saturate(x); return x * x * (3.0f - x * 2.0f); }                      // ml:121
inline float4 SmoothStep01(float4 v) { return float4(SmoothStep01(v.x), SmoothStep01(v.y), SmoothStep01(v.z), SmoothStep01(v.w)); }
inline float SmoothStep(float a, float b, float x) { x = LinearStep(a, b, x); return x * x * (3.0f - x * 2.0f); }  // ml:133
inline float Sign(float x) { return sign_fast(x); }                                                           // ml:168
inline float Pow01(float x, float y) { return std::pow(saturate(x), y); }                                      // ml:203
inline float Sqrt01(float x) { return std::sqrt(saturate(x)); }                                                // ml:241
inline float Rsqrt(float x) { return 1.0f / std::sqrt(max(x, ML_SMALL_EPS)); }                                 // ml:259 (ACCURATE_SAFE)
inline float AcosApproxPositive(float x) { return lerp(1.567589f, 1.399331f, saturate(x)) * std::sqrt(saturate(1.0f - x)); }  // ml:298
inline float PositiveRcp(float x) { return 1.0f / max(x, ML_SMALL_EPS); }                                      // ml:357 (ACCURATE_SAFE)
inline float LengthSquared(float2 v) { return dot(v, v); }
inline float LengthSquared(float3 v) { return dot(v, v); }
}  // namespace Math

// ------------------------------------------------------------------------------------------------------------
// Geometry (ml:477-672)
// ------------------------------------------------------------------------------------------------------------
namespace Geometry {
inline float4 GetRotator(float angle) { float ca = std::cos(angle), sa = std::sin(angle); return float4(ca, sa, -sa, ca); }  // ml:479
inline float4 CombineRotators(float4 r1, float4 r2) {  // ml:513: r1.xyxy * r2.xxzz + r1.zwzw * r2.yyww
    return float4(r1.x * r2.x + r1.z * r2.y, r1.y * r2.x + r1.w * r2.y, r1.x * r2.z + r1.z * r2.w, r1.y * r2.z + r1.w * r2.w);
}
inline float4 ScaleRotator(float4 r, float2 scale) { return float4(scale.x * r.x, scale.x * r.z, scale.y * r.y, scale.y * r.w); }  // ml:519
inline float2 RotateVector(float4 rotator, float2 v) { return float2(v.x * rotator.x + v.y * rotator.y, v.x * rotator.z + v.y * rotator.w); }  // ml:522
inline float3 RotateVector(const float4x4& m, float3 v) { return mul3x3(m, v); }          // ml:526
inline float3 RotateVectorInverse(const float4x4& m, float3 v) { return mul3x3T(m, v); }  // ml:537
inline float3 AffineTransform(const float4x4& m, float3 p) { return mul(m, float4(p, 1.0f)).xyz(); }  // ml:544
inline float4 ProjectiveTransform(const float4x4& m, float3 p) { return mul(m, float4(p, 1.0f)); }    // ml:555

struct Basis { float3 T, B, N; };
inline Basis GetBasis(float3 N) {  // ml:574-588
    float sz = Math::Sign(N.z);
    float a = 1.0f / (sz + N.z);
    float ya = N.y * a;
    float b = N.x * ya;
    float c = N.x * sz;
    Basis r;
    r.T = float3(c * N.x * a - 1.0f, sz * b, c);
    r.B = float3(b, N.y * ya - sz, N.y);
    r.N = N;
    return r;
}
// mul( float3x3( T, B, N ), v ): rows are T, B, N
inline float3 RotateVector(const Basis& m, float3 v) { return float3(dot(m.T, v), dot(m.B, v), dot(m.N, v)); }

inline float3 ReconstructViewPosition(float2 uv, float4 frustum, float viewZ = 1.0f, float orthoMode = 0.0f) {  // ml:649
    float3 p;
    p.x = uv.x * frustum.z + frustum.x;
    p.y = uv.y * frustum.w + frustum.y;
    float s = orthoMode == 0.0f ? viewZ : orthoMode;
    p.x *= s;
    p.y *= s;
    p.z = viewZ;
    return p;
}
inline float2 GetScreenUv(const float4x4& worldToClip, float3 X) {  // ml:659, D3D window origin
    float4 clip = ProjectiveTransform(worldToClip, X);
    return float2(clip.x / clip.w, clip.y / clip.w) * float2(0.5f, -0.5f) + 0.5f;
}
}  // namespace Geometry

// ------------------------------------------------------------------------------------------------------------
// Color, Packing, Filtering, Sequence, Rng, ImportanceSampling
// ------------------------------------------------------------------------------------------------------------
namespace Color {
inline float Luminance(float3 x) { return dot(x, float3(0.2126f, 0.7152f, 0.0722f)); }  // ml:712-717
inline float Clamp(float m1, float sigma, float c) { return clamp(c, m1 - sigma, m1 + sigma); }  // ml:1101
}

namespace Packing {
// ml:1179-1212, LSB-first R|G|B|A
inline uint32_t RgbaToUint(float4 c, uint32_t Rbits, uint32_t Gbits, uint32_t Bbits, uint32_t Abits) {
    uint32_t mask[4] = {(1u << Rbits) - 1u, (1u << Gbits) - 1u, (1u << Bbits) - 1u, (1u << Abits) - 1u};
    uint32_t shift[4] = {0, Rbits, Rbits + Gbits, Rbits + Gbits + Bbits};
    uint32_t p = 0;
    for (int i = 0; i < 4; i++) p |= (uint32_t)(saturate(c[i]) * float(mask[i]) + 0.5f) << shift[i];
    return p;
}
inline float4 UintToRgba(uint32_t p, uint32_t Rbits, uint32_t Gbits, uint32_t Bbits, uint32_t Abits) {
    uint32_t mask[4] = {(1u << Rbits) - 1u, (1u << Gbits) - 1u, (1u << Bbits) - 1u, (1u << Abits) - 1u};
    uint32_t shift[4] = {0, Rbits, Rbits + Gbits, Rbits + Gbits + Bbits};
    float4 r;
    for (int i = 0; i < 4; i++) r[i] = float((p >> shift[i]) & mask[i]) * (1.0f / max(float(mask[i]), 1.0f));
    return r;
}
}  // namespace Packing

namespace Filtering {
inline float GetModifiedRoughnessFromNormalVariance(float linearRoughness, float3 nonNormalizedAverageNormal) {  // ml:1287
    float l = length(nonNormalizedAverageNormal);
    float kappa = saturate(1.0f - l * l) * Math::PositiveRcp(l * (3.0f - l * l));
    return Math::Sqrt01(linearRoughness * linearRoughness + kappa);
}
struct Bilinear { float2 origin, weights; };
inline Bilinear GetBilinearFilter(float2 uv, float2 texSize) {  // ml:1338
    float2 t = uv * texSize - 0.5f;
    Bilinear r;
    r.origin = floor(t);
    r.weights = saturate(t - r.origin);
    return r;
}
template <class T> inline T ApplyBilinearFilter(T s00, T s10, T s01, T s11, Bilinear f) {  // ml:1349
    return lerp(lerp(s00, s10, f.weights.x), lerp(s01, s11, f.weights.x), f.weights.y);
}
inline float4 GetBilinearCustomWeights(Bilinear f, float4 customWeights) {  // ml:1361
    float2 o = saturate(1.0f - f.weights);
    float4 w = customWeights;
    w.x *= o.x * o.y;
    w.y *= f.weights.x * o.y;
    w.z *= o.x * f.weights.y;
    w.w *= f.weights.x * f.weights.y;
    return w;
}
inline float ApplyBilinearCustomWeights(float s00, float s10, float s01, float s11, float4 w) {  // ml:1375, normalize = true
    float sum = dot(w, float4(1.0f));
    return (s00 * w.x + s10 * w.y + s01 * w.z + s11 * w.w) * (sum < 0.0001f ? 0.0f : rcp(sum));
}
struct CatmullRom { float2 origin; };
inline CatmullRom GetCatmullRomFilter(float2 uv, float2 texSize) {  // ml:1397 (only the origin is consumed by REBLUR)
    float2 tci = uv * texSize;
    float2 tc = floor(tci - 0.5f) + 0.5f;
    CatmullRom r;
    r.origin = tc - 1.5f;
    return r;
}
}  // namespace Filtering

namespace Sequence {
inline uint32_t CheckerBoard(uint32_t x, uint32_t y, uint32_t frameIndex) { return ((x ^ y) ^ frameIndex) & 1u; }  // ml:1620
inline uint32_t IntegerExplode(uint32_t x) {  // ml:1627
    x = (x | (x << 8)) & 0x00FF00FFu;
    x = (x | (x << 4)) & 0x0F0F0F0Fu;
    x = (x | (x << 2)) & 0x33333333u;
    x = (x | (x << 1)) & 0x55555555u;
    return x;
}
inline uint32_t Zorder(uint32_t x, uint32_t y) { return IntegerExplode(x) | (IntegerExplode(y) << 1); }  // ml:1638
inline uint32_t Hash(uint32_t x) {  // ml:1657
    x ^= x >> 16;
    x *= 0x7FEB352Du;
    x ^= x >> 15;
    x *= 0x846CA68Bu;
    x ^= x >> 16;
    return x;
}
inline uint32_t HashCombine(uint32_t seed, uint32_t value) { return seed ^ (Hash(value) + 0x9E3779B9u + (seed << 6) + (seed >> 2)); }  // ml:1679
}  // namespace Sequence

// Rng::Hash (ml:1834-1862), ML_RNG_NEXT_MODE = HASH, ML_RNG_FLOAT01_MODE = MANTISSA_BITS
struct RngHash {
    uint32_t state = 0;
    void Initialize(uint32_t x, uint32_t y, uint32_t frameIndex) {
        state = Sequence::HashCombine(Sequence::Hash(frameIndex + 0x035F9F29u), Sequence::Zorder(x, y));
    }
    uint32_t GetUint() { state = Sequence::Hash(state); return state; }
    float GetFloat() { uint32_t x = GetUint(); return asfloat((x >> 9) | 0x3F800000u) - 1.0f; }
    float2 GetFloat2() { float a = GetFloat(); float b = GetFloat(); return float2(a, b); }
};

namespace ImportanceSampling {
inline float GetSpecularLobeTanHalfAngle(float linearRoughness, float percentOfVolume = 0.75f) {  // ml:2347
    percentOfVolume = saturate(percentOfVolume);
    return saturate(linearRoughness) * std::sqrt(percentOfVolume / (1.0f - percentOfVolume + ML_EPS));
}
inline float GetSpecularDominantFactorG2(float NoV, float linearRoughness) {  // ml:2370, mode G2
    linearRoughness = saturate(linearRoughness);
    float a = 0.298475f * std::log(39.4115f - 39.0029f * linearRoughness);
    float dominantFactor = Math::Pow01(1.0f - NoV, 10.8649f) * (1.0f - a) + a;
    return saturate(dominantFactor);
}
inline float4 GetSpecularDominantDirectionG2(float3 N, float3 V, float linearRoughness) {  // ml:2399
    float NoV = std::fabs(dot(N, V));
    float dominantFactor = GetSpecularDominantFactorG2(NoV, linearRoughness);
    float3 R = reflect(-V, N);
    float3 D = lerp(N, R, dominantFactor);
    return float4(normalize(D), dominantFactor);
}
}  // namespace ImportanceSampling

// ------------------------------------------------------------------------------------------------------------
// NRD.hlsli
// ------------------------------------------------------------------------------------------------------------
inline float3 _NRD_SafeNormalize(float3 v) { return v * rsqrt(dot(v, v) + 1e-9f); }  // nrd:361
inline float4 _NRD_DecodeNormalRoughness101010(float3 p) {  // nrd:387
    float t = p.z * 2.0f - 1.0f;
    float4 r;
    r.x = p.x - p.y;
    r.y = p.x + p.y - 1.0f;
    r.z = t < 0.0f ? -1.0f : 1.0f;
    r.z *= 1.0f - std::fabs(r.x) - std::fabs(r.y);
    r.w = std::fabs(t);
    return r;
}
inline float3 _NRD_EncodeNormalRoughness101010(float3 n, float roughness) {  // nrd:370
    n = n / (std::fabs(n.x) + std::fabs(n.y) + std::fabs(n.z));
    float3 r;
    r.y = n.y * 0.5f + 0.5f;
    r.x = n.x * 0.5f + r.y;
    r.y -= n.x * 0.5f;
    roughness = max(roughness, 1.5f / 512.0f);
    float s = n.z < 0.0f ? -roughness : roughness;
    r.z = s * 0.5f + 0.5f;
    return r;
}
inline float3 _NRD_LinearToYCoCg(float3 c) {  // nrd:409
    return float3(dot(c, float3(0.25f, 0.5f, 0.25f)), dot(c, float3(0.5f, 0.0f, -0.5f)), dot(c, float3(-0.25f, 0.5f, -0.25f)));
}
inline float3 _NRD_YCoCgToLinear(float3 c) {  // nrd:418
    float t = c.x - c.z;
    float3 r;
    r.y = c.x + c.z;
    r.x = t + c.y;
    r.z = t - c.y;
    return max(r, float3(0.0f));
}
inline float _NRD_GetSpecMagicCurve(float roughness, float power) {  // nrd:559
    float f = 1.0f - std::exp2(-200.0f * roughness * roughness);
    f *= std::pow(saturate(roughness), power);
    return f;
}
inline float _REBLUR_GetHitDistanceNormalization(float viewZ, float3 hitDistParams, float roughness) {  // nrd:568
    float smc = _NRD_GetSpecMagicCurve(roughness, 0.5f);
    return (hitDistParams.x + std::fabs(viewZ) * hitDistParams.y) * lerp(hitDistParams.z, 1.0f, smc);
}
inline float4 NRD_FrontEnd_UnpackNormalAndRoughness(float4 p, float& materialID) {  // nrd:656
    float4 r = _NRD_DecodeNormalRoughness101010(p.xyz());
    materialID = p.w * 3.0f;
    float3 n = _NRD_SafeNormalize(r.xyz());
    return float4(n, r.w);
}
inline float4 NRD_FrontEnd_UnpackNormalAndRoughness(float4 p) { float unused; return NRD_FrontEnd_UnpackNormalAndRoughness(p, unused); }
inline float4 NRD_FrontEnd_UnpackRoughness(float4 r) { return abs(r * 2.0f - 1.0f); }  // nrd:637 (applied to gathered blue channel)
inline float4 NRD_FrontEnd_PackNormalAndRoughness(float3 N, float roughness, float materialID) {  // nrd:696
    float3 p = _NRD_EncodeNormalRoughness101010(N, roughness);
    return float4(p, saturate(materialID / 3.0f));
}
inline float REBLUR_FrontEnd_GetNormHitDist(float hitDist, float viewZ, float3 hitDistParams, float roughness) {  // nrd:806
    return saturate(hitDist / _REBLUR_GetHitDistanceNormalization(viewZ, hitDistParams, roughness));
}
inline float NRD_GetNormalizedStrandThickness(float strandThickness, float pixelSize) { return saturate(0.5f * pixelSize / (strandThickness + NRD_EPS)); }  // nrd:1303

// ------------------------------------------------------------------------------------------------------------
// Common.hlsli (functions that do not touch the constant buffer)
// ------------------------------------------------------------------------------------------------------------
static const float3 g_Special8[8] = {  // common:207-218
    float3(-1.0f, 0.0f, 1.0f), float3(0.0f, 1.0f, 1.0f), float3(1.0f, 0.0f, 1.0f), float3(0.0f, -1.0f, 1.0f),
    float3(-0.25f * 1.41421356237309504880f, 0.25f * 1.41421356237309504880f, 0.5f), float3(0.25f * 1.41421356237309504880f, 0.25f * 1.41421356237309504880f, 0.5f),
    float3(0.25f * 1.41421356237309504880f, -0.25f * 1.41421356237309504880f, 0.5f), float3(-0.25f * 1.41421356237309504880f, -0.25f * 1.41421356237309504880f, 0.5f)};

inline float GetStdDev(float m1, float m2) { return std::sqrt(std::fabs(m2 - m1 * m1)); }                         // common:253
inline bool CompareMaterials(float m0, float m, float minm) { return max(m0, minm) == max(m, minm); }             // common:256
inline float PixelRadiusToWorld(float unproject, float orthoMode, float pixelRadius, float viewZ) {               // common:264
    return pixelRadius * unproject * lerp(viewZ, 1.0f, std::fabs(orthoMode));
}
inline float GetFrustumSize(float minRectDimMulUnproject, float orthoMode, float viewZ) {                          // common:269
    return minRectDimMulUnproject * lerp(viewZ, 1.0f, std::fabs(orthoMode));
}
inline float GetHitDistFactor(float hitDist, float frustumSize) { return saturate(hitDist / frustumSize); }        // common:277
inline float IsInScreenNearest(float2 uv) { return float(uv.x > 0.0f && uv.y > 0.0f && uv.x < 1.0f && uv.y < 1.0f); }  // common:307
inline float2 MirrorUv(float2 uv) {                                                                                // common:312
    float2 m = 1.0f - abs(1.0f - frac(uv * 0.5f) * 2.0f);
    return min(m, float2(0.99999f));
}
inline float4 IsInScreenBilinear(float2 footprintOrigin, float2 rectSize) {                                        // common:322
    float4 p = float4(footprintOrigin.x, footprintOrigin.y, footprintOrigin.x + 1.0f, footprintOrigin.y + 1.0f);
    float4 r = float4(float(p.x >= 0.0f), float(p.y >= 0.0f), float(p.z >= 0.0f), float(p.w >= 0.0f));
    r = r * float4(float(p.x < rectSize.x), float(p.y < rectSize.y), float(p.z < rectSize.x), float(p.w < rectSize.y));
    return float4(r.x * r.y, r.z * r.y, r.x * r.w, r.z * r.w);  // r.xzxz * r.yyww
}
inline float GetSpecMagicCurve(float roughness, float power = 0.25f) { return _NRD_GetSpecMagicCurve(roughness, power); }  // common:346
inline float ComputeParallaxInPixels(float3 X, float2 uvForZeroParallax, const float4x4& mWorldToClip, float2 rectSize) {  // common:351
    float2 uv = Geometry::GetScreenUv(mWorldToClip, X);
    float2 parallaxInUv = uv - uvForZeroParallax;
    return length(parallaxInUv * rectSize);
}
inline float3 GetXvirtual(float hitDist, float curvature, float3 X, float3 Xprev, float3 N, float3 V, float roughness) {  // common:421
    float4 D = ImportanceSampling::GetSpecularDominantDirectionG2(N, V, roughness);
    float3 reflectionRay = D.xyz() * hitDist;
    Geometry::Basis reflectorBasis = Geometry::GetBasis(N);
    float3 O = Geometry::RotateVector(reflectorBasis, reflectionRay);
    O.z = -O.z;
    float mag = 1.0f / (2.0f * curvature * O.z - 1.0f);
    float NoV = std::fabs(dot(N, V));
    float f = length(X);
    f *= saturate(1.0f - NoV);
    f *= max(curvature, 0.0f);
    f = 1.0f / (1.0f + f);
    mag *= f;
    float3 I = O * mag;
    D.w *= length(I);
    float closenessToSurface = saturate(D.w / (hitDist + NRD_EPS));
    float3 x = lerp(Xprev, X, closenessToSurface);
    return x + V * D.w * Math::Sign(mag);
}
inline float2 GetKernelSampleCoordinates(const float4x4& mToClip, float3 offset, float3 X, float3 T, float3 B, float4 rotator) {  // common:463
    float2 o = Geometry::RotateVector(rotator, offset.xy());
    float3 p = X + T * o.x + B * o.y;
    float4 c4 = Geometry::ProjectiveTransform(mToClip, p);
    float3 clip = float3(c4.x, c4.y, c4.w);
    clip.x /= clip.z;
    clip.y /= clip.z;
    clip.y = -clip.y;
    return float2(clip.x, clip.y) * 0.5f + 0.5f;
}
inline float GetNormalWeightParam(float nonLinearAccumSpeed, float lobeAngleFraction, float roughness = 1.0f) {  // common:484
    float percentOfVolume = NRD_MAX_PERCENT_OF_LOBE_VOLUME * lerp(saturate(lobeAngleFraction), 1.0f, nonLinearAccumSpeed);
    float tanHalfAngle = ImportanceSampling::GetSpecularLobeTanHalfAngle(roughness, percentOfVolume);
    float angle = max(std::atan(tanHalfAngle), NRD_NORMAL_ENCODING_ERROR);
    return 1.0f / angle;
}
inline float2 GetGeometryWeightParams(float planeDistSensitivity, float frustumSize, float3 Xv, float3 Nv) {  // common:498
    float norm = planeDistSensitivity * frustumSize;
    float a = 1.0f / norm;
    float b = dot(Nv, Xv) * a;
    return float2(a, -b);
}
inline float2 GetHitDistanceWeightParams(float hitDist, float nonLinearAccumSpeed) {  // common:508
    float a = 1.0f / nonLinearAccumSpeed;
    float b = hitDist * a;
    return float2(a, -b);
}
inline float2 GetRoughnessWeightParams(float roughness, float fraction, float sensitivity = NRD_ROUGHNESS_SENSITIVITY) {  // common:519
    float a = 1.0f / lerp(sensitivity, 1.0f, saturate(roughness * fraction));
    float b = roughness * a;
    return float2(a, -b);
}
inline float2 GetRelaxedRoughnessWeightParams(float m, float fraction = 1.0f, float sensitivity = NRD_ROUGHNESS_SENSITIVITY) {  // common:527
    float a = 1.0f / lerp(sensitivity, 1.0f, lerp(m * m, m, saturate(fraction)));
    float b = m * a;
    return float2(a, -b);
}
inline float ExpApprox(float x) { return rcp(x * x - x + 1.0f); }                                                  // common:544
inline float ComputeExponentialWeight(float x, float px, float py) { return ExpApprox(-3.0f * std::fabs(x * px + py)); }  // common:550
inline float ComputeNonExponentialWeight(float x, float px, float py) { return Math::SmoothStep(1.0f, 0.0f, std::fabs(x * px + py)); }  // common:555
inline float ComputeWeight(float x, float px, float py) { return ComputeNonExponentialWeight(x, px, py); }          // common:564 (NRD_USE_EXPONENTIAL_WEIGHTS = 0)
inline float GetGaussianWeight(float r) { return std::exp(-0.66f * r * r); }                                       // common:574
inline float GetEncodingAwareNormalWeight(float3 Ncurr, float3 Nprev, float maxAngle, float curvatureAngle, float thresholdAngle) {  // common:581
    float cosa = dot(Ncurr, Nprev);
    float angle = Math::AcosApproxPositive(cosa);
    float w = Math::SmoothStep01(1.0f - (angle - curvatureAngle - thresholdAngle) / maxAngle);
    w = Math::SmoothStep(0.05f, 0.95f, w);
    return w;
}
const float NRD_DISOCCLUSION_THRESHOLD = 0.02f;  // common:63
inline float GetDisocclusionThreshold(float disocclusionThreshold, float frustumSize, float NoV) {  // common:595
    return frustumSize * saturate(disocclusionThreshold / max(0.05f, NoV));
}

}  // namespace orc
