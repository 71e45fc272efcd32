/* nrd_frontend.cuh — the application side of the denoiser, for CUDA / OptiX renderers.
 *
 * What it replaces: the front-end / back-end helpers of External/NRD/Shaders/NRD.hlsli (the HLSL header an application includes to PACK
 * its ray-tracing results into the denoiser's input textures and to UNPACK / RESOLVE the denoised outputs), i.e. the step either side
 * of the pass chains: Shaders/TraceOpaque.cs.hlsl:738-757 (pack) and Shaders/Composition.cs.hlsl:85-118 (unpack + SH / SG resolve) in
 * NRDSample. Every function cites the NRD.hlsli lines it restates; names and argument order are the reference's.
 *
 * Header-only, `__host__ __device__`: the same text runs in a CUDA kernel and in host C++ (tests/test_frontend_codecs.py compares the
 * host build bit for bit with NRD.hlsli itself compiled as C++, and the device build through nrdcuFrontEnd* in nrdcu.h).
 * Encodings are the ones the library is built with (NRD/CMakeLists.txt:77-88): NRD_NORMAL_ENCODING = R10G10B10A2_UNORM ( 2 ),
 * NRD_ROUGHNESS_ENCODING = LINEAR ( 1 ). Plain fp32, no fast-math assumptions: `a * b + c` is written as the reference writes it.
 */
#ifndef NRD_FRONTEND_CUH
#define NRD_FRONTEND_CUH

#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#    define NRDFE_FN __host__ __device__ inline
#else
#    define NRDFE_FN inline
#endif

namespace nrdfe {

// ---------------------------------------------------------------------------------------------------------------------------------
// Small vector PODs (kept private to this header so that it also compiles without cuda_runtime.h)
// ---------------------------------------------------------------------------------------------------------------------------------
struct F2 { float x, y; };
struct F3 { float x, y, z; };
struct F4 { float x, y, z, w; };

NRDFE_FN F2 f2(float x, float y) { F2 r = {x, y}; return r; }
NRDFE_FN F3 f3(float x, float y, float z) { F3 r = {x, y, z}; return r; }
NRDFE_FN F4 f4(float x, float y, float z, float w) { F4 r = {x, y, z, w}; return r; }
NRDFE_FN F4 f4(F3 v, float w) { F4 r = {v.x, v.y, v.z, w}; return r; }
NRDFE_FN F3 xyz(F4 v) { return f3(v.x, v.y, v.z); }
NRDFE_FN F3 operator+(F3 a, F3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
NRDFE_FN F3 operator-(F3 a, F3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
NRDFE_FN F3 operator-(F3 a) { return f3(-a.x, -a.y, -a.z); }
NRDFE_FN F3 operator*(F3 a, F3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
NRDFE_FN F3 operator*(F3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
NRDFE_FN F3 operator*(float s, F3 a) { return f3(s * a.x, s * a.y, s * a.z); }
NRDFE_FN F3 operator/(F3 a, float s) { return f3(a.x / s, a.y / s, a.z / s); }
NRDFE_FN F2 operator+(F2 a, F2 b) { return f2(a.x + b.x, a.y + b.y); }
NRDFE_FN F2 operator*(F2 a, float s) { return f2(a.x * s, a.y * s); }
NRDFE_FN float dot(F2 a, F2 b) { return a.x * b.x + a.y * b.y; }
NRDFE_FN float dot(F3 a, F3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
NRDFE_FN float dot(F4 a, F4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
NRDFE_FN float length(F3 a) { return sqrtf(dot(a, a)); }
NRDFE_FN F3 normalize(F3 a) { return a * (1.0f / length(a)); }
NRDFE_FN F3 reflect(F3 i, F3 n) { return i - 2.0f * dot(i, n) * n; }
NRDFE_FN float saturate(float x) { return fminf(fmaxf(x, 0.0f), 1.0f); }
NRDFE_FN F3 saturate(F3 v) { return f3(saturate(v.x), saturate(v.y), saturate(v.z)); }
NRDFE_FN float lerp(float a, float b, float t) { return a + (b - a) * t; }
NRDFE_FN F3 lerp(F3 a, F3 b, float t) { return f3(lerp(a.x, b.x, t), lerp(a.y, b.y, t), lerp(a.z, b.z, t)); }
NRDFE_FN F3 lerp(F3 a, F3 b, F3 t) { return f3(lerp(a.x, b.x, t.x), lerp(a.y, b.y, t.y), lerp(a.z, b.z, t.z)); }
NRDFE_FN F3 max3(F3 a, float b) { return f3(fmaxf(a.x, b), fmaxf(a.y, b), fmaxf(a.z, b)); }
NRDFE_FN F3 clamp3(F3 a, float lo, float hi) { return f3(fminf(fmaxf(a.x, lo), hi), fminf(fmaxf(a.y, lo), hi), fminf(fmaxf(a.z, lo), hi)); }
NRDFE_FN bool isInvalid(float x) { return isnan(x) || isinf(x); }                                   // _NRD_IsInvalid, NRD.hlsli:577
NRDFE_FN bool isInvalid(F3 v) { return isInvalid(v.x) || isInvalid(v.y) || isInvalid(v.z); }

// Constants (NRD.hlsli:93-106, 354-358)
#define NRDFE_FP16_MAX 65504.0f
#define NRDFE_PI 3.14159265358979323846f
#define NRDFE_EPS 1e-6f
#define NRDFE_INF 1e6f
#define NRDFE_REJITTER_VIEWZ_THRESHOLD 0.01f
#define NRDFE_REJITTER_AMPLITUDE 2.0f
#define NRDFE_MATERIAL_FACTOR_MIN_SCALE 0.02f
#define NRDFE_ROUGHNESS_FACTOR_MIN_SCALE 0.1f

// ---------------------------------------------------------------------------------------------------------------------------------
// Private helpers (NRD.hlsli:361-573)
// ---------------------------------------------------------------------------------------------------------------------------------
NRDFE_FN F3 _NRD_SafeNormalize(F3 v) { return v * (1.0f / sqrtf(dot(v, v) + 1e-9f)); }               // :361-364 ( rsqrt )

NRDFE_FN F3 _NRD_EncodeNormalRoughness101010(F3 n, float roughness) {                               // :370-385
    n = n / (fabsf(n.x) + fabsf(n.y) + fabsf(n.z));
    F3 r;
    r.y = n.y * 0.5f + 0.5f;
    r.x = n.x * 0.5f + r.y;
    r.y -= n.x * 0.5f;
    roughness = fmaxf(roughness, 1.5f / 512.0f);  // the sign of the stored value carries sign( n.z ): keep it away from +-0
    float s = n.z < 0.0f ? -roughness : roughness;
    r.z = s * 0.5f + 0.5f;
    return r;
}
NRDFE_FN F4 _NRD_DecodeNormalRoughness101010(F3 p) {                                                // :387-400
    float t = p.z * 2.0f - 1.0f;  // [ -1, 1 ]: magnitude = roughness, sign = sign( n.z )
    F4 r;
    r.x = p.x - p.y;
    r.y = p.x + p.y - 1.0f;
    r.z = t < 0.0f ? -1.0f : 1.0f;
    r.z *= 1.0f - fabsf(r.x) - fabsf(r.y);
    r.w = fabsf(t);
    return r;  // un-normalised; the caller normalises
}
NRDFE_FN float _NRD_Luminance(F3 c) { return dot(c, f3(0.2126f, 0.7152f, 0.0722f)); }                // :403-406
NRDFE_FN F3 _NRD_LinearToYCoCg(F3 c) {                                                              // :409-416
    return f3(dot(c, f3(0.25f, 0.5f, 0.25f)), dot(c, f3(0.5f, 0.0f, -0.5f)), dot(c, f3(-0.25f, 0.5f, -0.25f)));
}
NRDFE_FN F3 _NRD_YCoCgToLinear(F3 c) {                                                              // :418-428
    float t = c.x - c.z;
    F3 r;
    r.y = c.x + c.z;
    r.x = t + c.y;
    r.z = t - c.y;
    return max3(r, 0.0f);
}
NRDFE_FN F3 _NRD_YCoCgToLinear_Corrected(float Y, float Y0, F2 CoCg) {                              // :430-436
    Y = fmaxf(Y, 0.0f);
    CoCg = CoCg * ((Y + NRDFE_EPS) / (Y0 + NRDFE_EPS));
    return _NRD_YCoCgToLinear(f3(Y, CoCg.x, CoCg.y));
}
NRDFE_FN float _NRD_GetSpecularDominantFactor(float NoV, float roughness) {                         // :439-445
    float a = 0.298475f * logf(39.4115f - 39.0029f * roughness);
    float dominantFactor = powf(saturate(1.0f - NoV), 10.8649f) * (1.0f - a) + a;
    return saturate(dominantFactor);
}
NRDFE_FN F3 _NRD_GetSpecularDominantDirection(F3 N, F3 V, float dominantFactor) {                   // :447-453
    F3 R = reflect(-V, N);
    F3 D = lerp(N, R, dominantFactor);
    return normalize(D);
}
NRDFE_FN float _NRD_Pow5(float x) { return powf(saturate(1.0f - x), 5.0f); }                         // :455-458
NRDFE_FN float _NRD_DistributionTerm(float roughness, float NoH) {                                  // :461-471
    float m = roughness * roughness;
    float m2 = m * m;
    float t = (NoH * m2 - NoH) * NoH + 1.0f;
    float a = m / t;
    float d = a * a;
    return d / NRDFE_PI;
}
NRDFE_FN float _NRD_GeometryTerm(float roughness, float NoL, float NoV) {                           // :474-483
    float m = roughness * roughness;
    float m2 = m * m;
    float a = NoV * sqrtf((NoL - m2 * NoL) * NoL + m2);
    float b = NoL * sqrtf((NoV - m2 * NoV) * NoV + m2);
    return 0.5f / (a + b);
}
NRDFE_FN float _NRD_DiffuseTerm(float roughness, float NoL, float NoV, float VoH) {                 // :486-494
    float f = 2.0f * VoH * VoH * roughness - 0.5f;  // Burley diffuse on LINEAR roughness, as the application front end expects
    float FdV = f * _NRD_Pow5(NoV) + 1.0f;
    float FdL = f * _NRD_Pow5(NoL) + 1.0f;
    float d = FdV * FdL;
    return d / NRDFE_PI;
}
NRDFE_FN F2 _NRD_ComputeBrdfs(F3 Ld, F3 Ls, F3 N, F3 V, float roughness) {                          // :497-526
    F2 result;
    float NoV = fabsf(dot(N, V));
    {  // lobe 0
        F3 H = normalize(Ld + V);
        float NoL = saturate(dot(N, Ld));
        float VoH = fabsf(dot(V, H));
        float Kdiff = _NRD_DiffuseTerm(roughness, NoL, NoV, VoH);
        result.x = Kdiff * NoL;
    }
    {  // lobe 1
        F3 H = normalize(Ls + V);
        float NoL = saturate(dot(N, Ls));
        float NoH = saturate(dot(N, H));
        float D = _NRD_DistributionTerm(roughness, NoH);
        float Gmod = _NRD_GeometryTerm(roughness, NoL, NoV);
        float Kspec = D * Gmod;
        result.y = Kspec * NoL;
    }
    return result;  // Fresnel left out: the material demodulation has divided it away
}
NRDFE_FN F3 _NRD_EnvironmentTerm_Rtg(F3 Rf0, float NoV, float roughness) {                          // :529-556 ( "Ray Tracing Gems", ch. 32 )
    float m = saturate(roughness * roughness);
    F4 X = f4(1.0f, NoV, NoV * NoV, 0.0f);
    X.w = NoV * X.z;
    F4 Y = f4(1.0f, m, m * m, 0.0f);
    Y.w = m * Y.z;
    // mul( M, v ) = ( dot( row0, v ), dot( row1, v ), .. ) with the row-major constructors of the reference
    F2 m1 = f2(dot(f2(0.99044f, -1.28514f), f2(X.x, X.y)), dot(f2(1.29678f, -0.755907f), f2(X.x, X.y)));
    F3 Xxyw = f3(X.x, X.y, X.w), Xxzw = f3(X.x, X.z, X.w), Yxyw = f3(Y.x, Y.y, Y.w);
    F3 m2 = f3(dot(f3(1.0f, 2.92338f, 59.4188f), Xxyw), dot(f3(20.3225f, -27.0302f, 222.592f), Xxyw), dot(f3(121.563f, 626.13f, 316.627f), Xxyw));
    F2 m3 = f2(dot(f2(0.0365463f, 3.32707f), f2(X.x, X.y)), dot(f2(9.0632f, -9.04756f), f2(X.x, X.y)));
    F3 m4 = f3(dot(f3(1.0f, 3.59685f, -1.36772f), Xxzw), dot(f3(9.04401f, -16.3174f, 9.22949f), Xxzw), dot(f3(5.56589f, 19.7886f, -20.2123f), Xxzw));
    float bias = dot(m1, f2(Y.x, Y.y)) * (1.0f / fmaxf(dot(m2, Yxyw), NRDFE_EPS));
    float scale = dot(m3, f2(Y.x, Y.y)) * (1.0f / fmaxf(dot(m4, Yxyw), NRDFE_EPS));
    return saturate(Rf0 * scale + f3(bias, bias, bias));
}
NRDFE_FN float _NRD_GetSpecMagicCurve(float roughness, float power = 0.25f) {                       // :559-566
    float f = 1.0f - exp2f(-200.0f * roughness * roughness);
    f *= powf(saturate(roughness), power);
    return f;
}
NRDFE_FN float _REBLUR_GetHitDistanceNormalization(float viewZ, F3 hitDistParams, float roughness) {  // :568-573
    float smc = _NRD_GetSpecMagicCurve(roughness, 0.5f);
    return (hitDistParams.x + fabsf(viewZ) * hitDistParams.y) * lerp(hitDistParams.z, 1.0f, smc);
}

// ---------------------------------------------------------------------------------------------------------------------------------
// Spherical Gaussian / SH carrier (NRD.hlsli:583-631)
// ---------------------------------------------------------------------------------------------------------------------------------
struct NRD_SG {
    float c0;
    F2 chroma;
    float normHitDist;
    F3 c1;
    float sharpness;
};
NRDFE_FN NRD_SG _NRD_SG_Create(F3 radiance, F3 direction, float normHitDist) {                      // :593-605
    F3 YCoCg = _NRD_LinearToYCoCg(radiance);
    NRD_SG sg;
    sg.c0 = YCoCg.x;
    sg.chroma = f2(YCoCg.y, YCoCg.z);
    sg.c1 = direction * YCoCg.x;
    sg.normHitDist = normHitDist;
    sg.sharpness = 0.0f;  // the resolve functions fill this in
    return sg;
}
NRDFE_FN float _NRD_SG_InnerProduct(NRD_SG a, NRD_SG b) {                                           // :619-631
    F3 dir = a.sharpness * a.c1 + b.sharpness * b.c1;
    float d = length(dir);
    float c = expf(d - a.sharpness - b.sharpness);
    c *= 1.0f - expf(-2.0f * d);
    c /= fmaxf(d, NRDFE_EPS);
    return 2.0f * NRDFE_PI * c * a.c0 * b.c0;
}

// ---------------------------------------------------------------------------------------------------------------------------------
// FRONT END: G-buffer (NRD.hlsli:637-731). `p` / the return values are the UNORM channel values of IN_NORMAL_ROUGHNESS; nrdfe::packR10G10B10A2
// turns them into the texel word.
// ---------------------------------------------------------------------------------------------------------------------------------
NRDFE_FN F4 NRD_FrontEnd_UnpackRoughness(F4 r) {                                                    // :637-652
    return f4(fabsf(r.x * 2.0f - 1.0f), fabsf(r.y * 2.0f - 1.0f), fabsf(r.z * 2.0f - 1.0f), fabsf(r.w * 2.0f - 1.0f));
}
NRDFE_FN F4 NRD_FrontEnd_UnpackNormalAndRoughness(F4 p, float& materialID) {                        // :656-684
    F4 r = _NRD_DecodeNormalRoughness101010(xyz(p));
    materialID = p.w * 3.0f;
    return f4(_NRD_SafeNormalize(xyz(r)), r.w);
}
NRDFE_FN F4 NRD_FrontEnd_UnpackNormalAndRoughness(F4 p) {                                           // :687-692
    float unused;
    return NRD_FrontEnd_UnpackNormalAndRoughness(p, unused);
}
NRDFE_FN F4 NRD_FrontEnd_PackNormalAndRoughness(F3 N, float roughness, float materialID) {          // :696-731
    return f4(_NRD_EncodeNormalRoughness101010(N, roughness), saturate(materialID / 3.0f));
}
// UNORM channel values <-> the R10G10B10A2_UNORM texel (round to nearest, like a typed store / load)
NRDFE_FN uint32_t packR10G10B10A2(F4 p) {
    uint32_t x = (uint32_t)(saturate(p.x) * 1023.0f + 0.5f), y = (uint32_t)(saturate(p.y) * 1023.0f + 0.5f), z = (uint32_t)(saturate(p.z) * 1023.0f + 0.5f);
    return x | (y << 10) | (z << 20) | ((uint32_t)(saturate(p.w) * 3.0f + 0.5f) << 30);
}
NRDFE_FN F4 unpackR10G10B10A2(uint32_t v) {
    return f4((float)(v & 1023u) / 1023.0f, (float)((v >> 10) & 1023u) / 1023.0f, (float)((v >> 20) & 1023u) / 1023.0f, (float)(v >> 30) / 3.0f);
}

// Material de-modulation factors (NRD.hlsli:735-751)
NRDFE_FN void NRD_MaterialFactors(F3 N, F3 V, F3 albedo, F3 Rf0, float roughness, F3& diffFactor, F3& specFactor) {
    const F3 one = f3(1.0f, 1.0f, 1.0f), minScale = f3(NRDFE_MATERIAL_FACTOR_MIN_SCALE, NRDFE_MATERIAL_FACTOR_MIN_SCALE, NRDFE_MATERIAL_FACTOR_MIN_SCALE);
    float NoV = fabsf(dot(N, V));
    F3 Fenv = _NRD_EnvironmentTerm_Rtg(Rf0, NoV, roughness);
    diffFactor = (one - Fenv) * albedo;
    diffFactor = lerp(minScale, one, diffFactor);
    specFactor = Fenv;
    specFactor = specFactor * lerp(f3(NRDFE_ROUGHNESS_FACTOR_MIN_SCALE, NRDFE_ROUGHNESS_FACTOR_MIN_SCALE, NRDFE_ROUGHNESS_FACTOR_MIN_SCALE), one, roughness);
    specFactor = lerp(minScale, one, specFactor);
}

// Hit distance averaging for many paths per pixel (NRD.hlsli:772-797)
#define NRDFE_INF_INTERNAL 3.40282347e+38f
NRDFE_FN float NRD_FrontEnd_SpecHitDistAveraging_Begin() { return NRDFE_INF_INTERNAL; }
NRDFE_FN float NRD_FrontEnd_TrimHitDistance(float hitDist, float threshold) { return hitDist < threshold ? 0.0f : hitDist; }
NRDFE_FN void NRD_FrontEnd_SpecHitDistAveraging_Add(float& accumulatedSpecHitDist, float hitDist) {
    accumulatedSpecHitDist = fminf(accumulatedSpecHitDist, hitDist == 0.0f ? NRDFE_INF_INTERNAL : hitDist);
}
NRDFE_FN void NRD_FrontEnd_SpecHitDistAveraging_End(float& accumulatedSpecHitDist) {
    accumulatedSpecHitDist = accumulatedSpecHitDist == NRDFE_INF_INTERNAL ? 0.0f : accumulatedSpecHitDist;
}

// ---------------------------------------------------------------------------------------------------------------------------------
// FRONT END / BACK END: REBLUR (NRD.hlsli:805-908)
// ---------------------------------------------------------------------------------------------------------------------------------
NRDFE_FN float REBLUR_FrontEnd_GetNormHitDist(float hitDist, float viewZ, F3 hitDistParams, float roughness = 1.0f) {  // :805-810
    float f = _REBLUR_GetHitDistanceNormalization(viewZ, hitDistParams, roughness);
    return saturate(hitDist / f);
}
NRDFE_FN F4 REBLUR_FrontEnd_PackRadianceAndNormHitDist(F3 radiance, float normHitDist, bool sanitize = true) {        // :815-826
    if (sanitize) {
        radiance = isInvalid(radiance) ? f3(0, 0, 0) : clamp3(radiance, 0.0f, NRDFE_FP16_MAX);
        normHitDist = isInvalid(normHitDist) ? 0.0f : saturate(normHitDist);
    }
    return f4(_NRD_LinearToYCoCg(radiance), normHitDist);
}
NRDFE_FN F4 REBLUR_FrontEnd_PackSh(F3 radiance, float normHitDist, F3 direction, F4& out1, bool sanitize = true) {    // :831-849
    if (sanitize) {
        radiance = isInvalid(radiance) ? f3(0, 0, 0) : clamp3(radiance, 0.0f, NRDFE_FP16_MAX);
        normHitDist = isInvalid(normHitDist) ? 0.0f : saturate(normHitDist);
        direction = isInvalid(direction) ? f3(0, 0, 0) : clamp3(direction, -1.0f, 1.0f);
    }
    NRD_SG sg = _NRD_SG_Create(radiance, direction, normHitDist);
    out1 = f4(sg.c1, sg.sharpness);
    return f4(sg.c0, sg.chroma.x, sg.chroma.y, sg.normHitDist);
}
NRDFE_FN F4 REBLUR_FrontEnd_PackDirectionalOcclusion(F3 direction, float normHitDist, bool sanitize = true) {         // :853-866
    if (sanitize) {
        direction = isInvalid(direction) ? f3(0, 0, 0) : clamp3(direction, -1.0f, 1.0f);
        normHitDist = isInvalid(normHitDist) ? 0.0f : saturate(normHitDist);
    }
    NRD_SG sg = _NRD_SG_Create(f3(normHitDist, normHitDist, normHitDist), direction, normHitDist);
    return f4(sg.c1, sg.c0);
}
NRDFE_FN F4 REBLUR_BackEnd_UnpackRadianceAndNormHitDist(F4 data) { return f4(_NRD_YCoCgToLinear(xyz(data)), data.w); }  // :870-875
NRDFE_FN NRD_SG REBLUR_BackEnd_UnpackSh(F4 sh0, F3 sh1) {                                                            // :879-890
    NRD_SG sg;
    sg.c0 = sh0.x;
    sg.chroma = f2(sh0.y, sh0.z);
    sg.normHitDist = sh0.w;
    sg.c1 = sh1;
    sg.sharpness = 0.0f;
    return sg;
}
NRDFE_FN NRD_SG REBLUR_BackEnd_UnpackDirectionalOcclusion(F4 data) {                                                 // :892-903
    NRD_SG sg;
    sg.c0 = data.w;
    sg.chroma = f2(0.0f, 0.0f);
    sg.normHitDist = data.w;
    sg.c1 = xyz(data);
    sg.sharpness = 0.0f;
    return sg;
}
NRDFE_FN float REBLUR_GetHitDist(float normHitDist, float viewZ, F3 hitDistParams, float roughness) {                // :1293-1298
    return normHitDist * _REBLUR_GetHitDistanceNormalization(viewZ, hitDistParams, roughness);
}

// ---------------------------------------------------------------------------------------------------------------------------------
// FRONT END / BACK END: RELAX (NRD.hlsli:912-966)
// ---------------------------------------------------------------------------------------------------------------------------------
NRDFE_FN F4 RELAX_FrontEnd_PackRadianceAndHitDist(F3 radiance, float hitDist, bool sanitize = true) {                 // :912-921
    if (sanitize) {
        radiance = isInvalid(radiance) ? f3(0, 0, 0) : clamp3(radiance, 0.0f, NRDFE_FP16_MAX);
        hitDist = isInvalid(hitDist) ? 0.0f : fminf(fmaxf(hitDist, 0.0f), NRDFE_FP16_MAX);
    }
    return f4(radiance, hitDist);
}
NRDFE_FN F4 RELAX_FrontEnd_PackSh(F3 radiance, float hitDist, F3 direction, F4& out1, bool sanitize = true) {        // :925-943
    if (sanitize) {
        radiance = isInvalid(radiance) ? f3(0, 0, 0) : clamp3(radiance, 0.0f, NRDFE_FP16_MAX);
        hitDist = isInvalid(hitDist) ? 0.0f : fminf(fmaxf(hitDist, 0.0f), NRDFE_FP16_MAX);
        direction = isInvalid(direction) ? f3(0, 0, 0) : clamp3(direction, -1.0f, 1.0f);
    }
    out1 = f4(direction * _NRD_Luminance(radiance), 0.0f);
    return f4(radiance, hitDist);
}
NRDFE_FN F4 RELAX_BackEnd_UnpackRadiance(F4 color) { return color; }                                                 // :947-950
NRDFE_FN NRD_SG RELAX_BackEnd_UnpackSh(F4 sh0, F3 sh1) { return REBLUR_BackEnd_UnpackSh(sh0, sh1); }                  // :954-965 ( same fields )

// ---------------------------------------------------------------------------------------------------------------------------------
// FRONT END / BACK END: SIGMA (NRD.hlsli:974-1010)
// ---------------------------------------------------------------------------------------------------------------------------------
NRDFE_FN float SIGMA_FrontEnd_PackPenumbra(float distanceToOccluder, float tanOfLightAngularRadius) {                // :974-981 ( infinite lights )
    float penumbraSize = distanceToOccluder * tanOfLightAngularRadius;
    float penumbraRadius = penumbraSize * 0.5f;
    return distanceToOccluder >= NRDFE_FP16_MAX ? NRDFE_FP16_MAX : fminf(penumbraRadius, 32768.0f);
}
NRDFE_FN float SIGMA_FrontEnd_PackPenumbra(float distanceToOccluder, float distanceToLight, float lightSize) {       // :985-992 ( local lights )
    float penumbraSize = lightSize * distanceToOccluder / fmaxf(distanceToLight - distanceToOccluder, NRDFE_EPS);
    float penumbraRadius = penumbraSize * 0.5f;
    return distanceToOccluder >= NRDFE_FP16_MAX ? NRDFE_FP16_MAX : fminf(penumbraRadius, 32768.0f);
}
NRDFE_FN F4 SIGMA_FrontEnd_PackTranslucency(float distanceToOccluder, F3 translucency) {                             // :994-1002
    F3 t = saturate(translucency);
    return f4(distanceToOccluder >= NRDFE_FP16_MAX ? 1.0f : 0.0f, t.x, t.y, t.z);
}
NRDFE_FN float SIGMA_BackEnd_UnpackShadow(float shadow) { return shadow * shadow; }                                  // :1010
NRDFE_FN F4 SIGMA_BackEnd_UnpackShadow(F4 shadow) { return f4(shadow.x * shadow.x, shadow.y * shadow.y, shadow.z * shadow.z, shadow.w * shadow.w); }

// ---------------------------------------------------------------------------------------------------------------------------------
// BACK END: SG / SH resolve (NRD.hlsli:1016-1194)
// ---------------------------------------------------------------------------------------------------------------------------------
NRDFE_FN F3 NRD_SG_ExtractColor(NRD_SG sg) { return _NRD_YCoCgToLinear(f3(sg.c0, sg.chroma.x, sg.chroma.y)); }       // :1016-1019
NRDFE_FN F3 NRD_SG_ExtractDirection(NRD_SG sg) { return sg.c1 / fmaxf(length(sg.c1), NRDFE_EPS); }                   // :1021-1024
// rows[ 3 ] = the rows of the float3x3 the reference passes to mul( rotation, sg.c1 )
NRDFE_FN void NRD_SG_Rotate(NRD_SG& sg, const F3 rows[3]) { sg.c1 = f3(dot(rows[0], sg.c1), dot(rows[1], sg.c1), dot(rows[2], sg.c1)); }  // :1026-1029
NRDFE_FN F3 NRD_SG_ResolveDiffuse(NRD_SG sg, F3 N, F3 V, float roughness) {                                          // :1032-1071
    F3 L = NRD_SG_ExtractDirection(sg);
    float NoL = saturate(dot(N, L));
    NRD_SG light = {};
    light.sharpness = 2.0f;
    light.c0 = sg.c0 * light.sharpness;  // amplitude scaled so that the lobe integrates to c0
    light.c1 = L;
    NRD_SG ndf = {};
    ndf.c0 = 1.0f;
    ndf.c1 = N;
    ndf.sharpness = 2.0f;
    float Y = _NRD_SG_InnerProduct(ndf, light);
    F3 H = normalize(L + V);
    float NoV = fabsf(dot(N, V));
    float VoH = fabsf(dot(V, H));
    float Kdiff = _NRD_DiffuseTerm(roughness, NoL, NoV, VoH);
    Y *= Kdiff;
    Y *= lerp(1.0f, lerp(1.5f, 0.6f, roughness), _NRD_Pow5(NoV));
    Y = fmaxf(Y, sg.c0 / NRDFE_PI);
    return _NRD_YCoCgToLinear_Corrected(Y, sg.c0, sg.chroma);
}
NRDFE_FN F3 NRD_SG_ResolveSpecular(NRD_SG sg, F3 N, F3 V, float roughness) {                                         // :1074-1122
    roughness = fmaxf(roughness, 0.05f);
    float m = roughness * roughness;
    float m2 = m * m;
    F3 L = NRD_SG_ExtractDirection(sg);
    float NoL = saturate(dot(N, L));
    F3 H = normalize(L + V);
    float NoV = fabsf(dot(N, V));
    float VoH = fabsf(dot(V, H));
    NoV = lerp(0.02f, 1.0f, NoV);  // grazing view angles would otherwise brighten the result
    NRD_SG light = {};
    light.sharpness = 2.0f / m2;
    light.c0 = sg.c0 * light.sharpness;  // amplitude scaled so that the lobe integrates to c0
    light.c1 = L;
    float ndfSharpness = 0.5f / fmaxf(m2 * VoH, 1e-8f);
    NRD_SG warpedNdf = {};
    warpedNdf.c0 = 1.0f;
    warpedNdf.c1 = L;  // = V mirrored about the half vector
    warpedNdf.sharpness = ndfSharpness;
    float Y = _NRD_SG_InnerProduct(warpedNdf, light);
    float Gmod = _NRD_GeometryTerm(roughness, NoL, NoV);
    Y *= Gmod * NoL;  // no Fresnel here, see _NRD_ComputeBrdfs
    Y *= lerp(lerp(0.1f, 0.4f, m2), 0.8f, NoV);
    Y = fmaxf(Y, sg.c0 / NRDFE_PI);
    return _NRD_YCoCgToLinear_Corrected(Y, sg.c0, sg.chroma);
}
// Offsets: e = ( 1, 0 ), w = ( -1, 0 ), n = ( 0, 1 ), s = ( 0, -1 )
NRDFE_FN F2 NRD_SG_ReJitter(NRD_SG diffSg, NRD_SG specSg, F3 V, float roughness, float Z, float Ze, float Zw, float Zn, float Zs, F3 N, F3 Ne, F3 Nw, F3 Nn,
                            F3 Ns) {                                                                              // :1131-1168
    F3 Ld = NRD_SG_ExtractDirection(diffSg);
    F3 Ls = NRD_SG_ExtractDirection(specSg);
    Ls = normalize(lerp(V, Ls, roughness));
    F2 brdfCenter = _NRD_ComputeBrdfs(Ld, Ls, N, V, roughness);
    F2 brdfAverage = _NRD_ComputeBrdfs(Ld, Ls, Ne, V, roughness);
    brdfAverage = brdfAverage + _NRD_ComputeBrdfs(Ld, Ls, Nn, V, roughness);
    brdfAverage = brdfAverage + _NRD_ComputeBrdfs(Ld, Ls, Nw, V, roughness);
    brdfAverage = brdfAverage + _NRD_ComputeBrdfs(Ld, Ls, Ns, V, roughness);
    brdfAverage = brdfAverage * 0.25f;
    F2 j = f2((brdfCenter.x + NRDFE_EPS) / (brdfAverage.x + NRDFE_EPS), (brdfCenter.y + NRDFE_EPS) / (brdfAverage.y + NRDFE_EPS));
    j = f2(fminf(fmaxf(j.x, 1.0f / NRDFE_REJITTER_AMPLITUDE), NRDFE_REJITTER_AMPLITUDE), fminf(fmaxf(j.y, 1.0f / NRDFE_REJITTER_AMPLITUDE), NRDFE_REJITTER_AMPLITUDE));
    float NoV = fabsf(dot(N, V));
    float zThreshold = NRDFE_REJITTER_VIEWZ_THRESHOLD * fabsf(Z) / (NoV * 0.95f + 0.05f);
    F4 w = f4(zThreshold >= fabsf(Ze - Z) ? 1.0f : 0.0f, zThreshold >= fabsf(Zw - Z) ? 1.0f : 0.0f, zThreshold >= fabsf(Zn - Z) ? 1.0f : 0.0f,
              zThreshold >= fabsf(Zs - Z) ? 1.0f : 0.0f);
    bool isSymmetrical = dot(w, f4(1.0f, 1.0f, 1.0f, 1.0f)) > 3.5f;
    return isSymmetrical ? j : f2(1.0f, 1.0f);
}
NRDFE_FN F3 NRD_SH_ResolveDiffuse(NRD_SG sh, F3 N) {                                                                  // :1172-1180
    const float k0 = 1.0f / NRDFE_PI, k1 = 3.0f / NRDFE_PI;
    float Y = sh.c0 * k0 + dot(sh.c1, N) * k1;
    return _NRD_YCoCgToLinear_Corrected(Y, sh.c0, sh.chroma);
}
NRDFE_FN F3 NRD_SH_ResolveSpecular(NRD_SG sh, F3 N, F3 V, float roughness) {                                          // :1182-1194
    const float k0 = 1.0f / NRDFE_PI, k1 = 3.0f / NRDFE_PI;
    float NoV = fabsf(dot(N, V));
    float f = _NRD_GetSpecularDominantFactor(NoV, roughness);
    F3 D = _NRD_GetSpecularDominantDirection(N, V, f);
    float Y = sh.c0 * k0 + dot(sh.c1, D) * k1;  // linear SH evaluation; NRD_SG_ResolveSpecular is the better estimator
    return _NRD_YCoCgToLinear_Corrected(Y, sh.c0, sh.chroma);
}

// ---------------------------------------------------------------------------------------------------------------------------------
// Misc (NRD.hlsli:1200-1308)
// ---------------------------------------------------------------------------------------------------------------------------------
NRDFE_FN float _NRD_SolidAngle(float coneCosAngle) { return 2.0f * NRDFE_PI * (1.0f - coneCosAngle); }                 // :1200-1203
NRDFE_FN float _NRD_AcosApproxSphere(float x) {                                                                      // :1205-1211
    float a = saturate(fabsf(x));
    float b = (0.5f * NRDFE_PI - 0.156583f * a) * sqrtf(1.0f - a);
    return x >= 0.0f ? b : (NRDFE_PI - b);
}
NRDFE_FN float _NRD_CosDifference(float cosA, float cosB) {                                                          // :1213-1219
    float sqSinA = saturate(1.0f - cosA * cosA);
    float sqSinB = saturate(1.0f - cosB * cosB);
    return cosA * cosB + sqrtf(sqSinA * sqSinB);  // angle-difference identity
}
NRDFE_FN F3 _NRD_RotateTowards(F3 a, F3 b, float cosAngle) {                                                         // :1221-1230
    float cosTheta = dot(a, b);
    float sinAngle = sqrtf(saturate(1.0f - cosAngle * cosAngle));
    float sinTheta = sqrtf(saturate(1.0f - cosTheta * cosTheta));
    float sinDiff = sinTheta * cosAngle - sinAngle * cosTheta;  // angle-difference identity
    F3 rotated = (sinDiff * a + sinAngle * b) * (1.0f / sinTheta);
    return sinTheta < NRDFE_EPS ? a : rotated;
}
NRDFE_FN F4 _NRD_GetSphericalCapIntersection(F3 axisA, float cosA, F3 axisB, float cosB) {                           // :1232-1266
    float radiusA = _NRD_AcosApproxSphere(cosA);
    float radiusB = _NRD_AcosApproxSphere(cosB);
    float cosDist = dot(axisA, axisB);
    float dist = _NRD_AcosApproxSphere(cosDist);
    if (dist > radiusA + radiusB) return f4(axisA, 1.0f);
    if (radiusB > radiusA + dist) return f4(axisA, cosA);
    if (radiusA > radiusB + dist) return f4(axisB, cosB);
    float diff = fabsf(radiusA - radiusB);
    float x = 1.0f - saturate((dist - diff) * (1.0f / (radiusA + radiusB - diff)));
    float area = x * x * (3.0f - 2.0f * x);
    float intersectionAngle = 1.0f - area * (1.0f - fmaxf(cosA, cosB));
    float cosDelta = _NRD_CosDifference(cosA, cosB);
    float cosAngle = sqrtf(0.5f * _NRD_CosDifference(cosDist, cosDelta) + 0.5f);
    F3 L = _NRD_RotateTowards(axisB, axisA, cosAngle);
    return f4(L, intersectionAngle);
}
NRDFE_FN float NRD_ComputeCavityShadow(NRD_SG sg, F3 N, float cavity, float cosLightAngle, float shadowStrength) {   // :1268-1281
    float coneCosAngle = sqrtf(saturate(1.0f - cavity));
    F3 lightDir = NRD_SG_ExtractDirection(sg);
    F4 coneIntersection = _NRD_GetSphericalCapIntersection(N, coneCosAngle, lightDir, cosLightAngle);
    float lightSolidAngle = _NRD_SolidAngle(cosLightAngle);
    float shadow = saturate(_NRD_SolidAngle(coneIntersection.w) / fmaxf(lightSolidAngle, NRDFE_EPS));
    return lerp(1.0f, shadow, shadowStrength);
}
NRDFE_FN bool NRD_IsValidRadiance(F3 radiance) { return !isInvalid(radiance); }                                      // :1285-1288
NRDFE_FN float NRD_GetNormalizedStrandThickness(float strandThickness, float pixelSize) {                           // :1303-1306
    return saturate(0.5f * pixelSize / (strandThickness + NRDFE_EPS));
}

}  // namespace nrdfe

#endif  // NRD_FRONTEND_CUH
