// TEST INFRASTRUCTURE — CPU oracle (see hlsl_like.h). The oracle's restatement of the MathLib helpers (nrd_shared.h),
// exported one by one so tests/test_oracle_math.py can compare each against the reference's own MathLib
// (oracle/_ref/libml_ref.so, built from /root/reference by ref_build.sh) and against tests/golden/ml_vectors.json.
#include "reblur_shared.h"

using namespace orc;
#define X extern "C" __attribute__((visibility("default")))

X float orc_LinearStep(float a, float b, float x) { return Math::LinearStep(a, b, x); }
X float orc_SmoothStep01(float x) { return Math::SmoothStep01(x); }
X float orc_SmoothStep(float a, float b, float x) { return Math::SmoothStep(a, b, x); }
X float orc_Pow01(float x, float y) { return Math::Pow01(x, y); }
X float orc_Sqrt01(float x) { return Math::Sqrt01(x); }
X float orc_AcosApproxPositive(float x) { return Math::AcosApproxPositive(x); }
X float orc_PositiveRcp(float x) { return Math::PositiveRcp(x); }
X float orc_Rsqrt(float x) { return Math::Rsqrt(x); }
X float orc_Sign(float x) { return Math::Sign(x); }
X void orc_GetRotator(float angle, float* o) { float4 r = Geometry::GetRotator(angle); o[0] = r.x; o[1] = r.y; o[2] = r.z; o[3] = r.w; }
X void orc_CombineRotators(const float* a, const float* b, float* o) {
    float4 r = Geometry::CombineRotators(float4(a[0], a[1], a[2], a[3]), float4(b[0], b[1], b[2], b[3]));
    o[0] = r.x; o[1] = r.y; o[2] = r.z; o[3] = r.w;
}
X void orc_ScaleRotator(const float* a, float sx, float sy, float* o) {
    float4 r = Geometry::ScaleRotator(float4(a[0], a[1], a[2], a[3]), float2(sx, sy));
    o[0] = r.x; o[1] = r.y; o[2] = r.z; o[3] = r.w;
}
X void orc_RotateVector2(const float* rot, float vx, float vy, float* o) {
    float2 r = Geometry::RotateVector(float4(rot[0], rot[1], rot[2], rot[3]), float2(vx, vy));
    o[0] = r.x; o[1] = r.y;
}
X void orc_ReconstructViewPosition(float u, float v, const float* frustum, float viewZ, float ortho, float* o) {
    float3 r = Geometry::ReconstructViewPosition(float2(u, v), float4(frustum[0], frustum[1], frustum[2], frustum[3]), viewZ, ortho);
    o[0] = r.x; o[1] = r.y; o[2] = r.z;
}
X void orc_GetScreenUv(const float* m16, const float* X3, float* o) {
    float4x4 m;
    for (int i = 0; i < 16; i++) m.m[i] = m16[i];
    float2 r = Geometry::GetScreenUv(m, float3(X3[0], X3[1], X3[2]));
    o[0] = r.x; o[1] = r.y;
}
X float orc_ColorClamp(float m1, float sigma, float c) { return Color::Clamp(m1, sigma, c); }
X uint32_t orc_RgbaToUint664(const float* c) { return Packing::RgbaToUint(float4(c[0], c[1], c[2], c[3]), 6, 6, 4, 0); }
X void orc_UintToRgba664(uint32_t p, float* o) { float4 r = Packing::UintToRgba(p, 6, 6, 4, 0); o[0] = r.x; o[1] = r.y; o[2] = r.z; o[3] = r.w; }
X float orc_GetModifiedRoughnessFromNormalVariance(float r, const float* n) { return Filtering::GetModifiedRoughnessFromNormalVariance(r, float3(n[0], n[1], n[2])); }
X void orc_GetBilinearFilter(float u, float v, float w, float h, float* o) {
    Filtering::Bilinear f = Filtering::GetBilinearFilter(float2(u, v), float2(w, h));
    o[0] = f.origin.x; o[1] = f.origin.y; o[2] = f.weights.x; o[3] = f.weights.y;
}
X void orc_GetBilinearCustomWeights(float ox, float oy, float wx, float wy, const float* c, float* o) {
    Filtering::Bilinear f; f.origin = float2(ox, oy); f.weights = float2(wx, wy);
    float4 r = Filtering::GetBilinearCustomWeights(f, float4(c[0], c[1], c[2], c[3]));
    o[0] = r.x; o[1] = r.y; o[2] = r.z; o[3] = r.w;
}
X float orc_ApplyBilinearCustomWeights(const float* s, const float* w) { return Filtering::ApplyBilinearCustomWeights(s[0], s[1], s[2], s[3], float4(w[0], w[1], w[2], w[3])); }
X void orc_GetCatmullRomOrigin(float u, float v, float w, float h, float* o) {
    Filtering::CatmullRom f = Filtering::GetCatmullRomFilter(float2(u, v), float2(w, h));
    o[0] = f.origin.x; o[1] = f.origin.y;
}
X uint32_t orc_Hash(uint32_t x) { return Sequence::Hash(x); }
X uint32_t orc_HashCombine(uint32_t s, uint32_t v) { return Sequence::HashCombine(s, v); }
X uint32_t orc_Zorder(uint32_t x, uint32_t y) { return Sequence::Zorder(x, y); }
X uint32_t orc_CheckerBoard(uint32_t x, uint32_t y, uint32_t f) { return Sequence::CheckerBoard(x, y, f); }
// host-side only (per-frame rotators): the C++ build of MathLib evaluates this in double
X float orc_Weyl1D(float p, uint32_t n) { double x = (double)p + (double)float(n * 10368889u) / 16777216.0; return (float)(x - std::floor(x)); }
X void orc_RngHash(uint32_t x, uint32_t y, uint32_t frame, float* o4) {
    RngHash r;
    r.Initialize(x, y, frame);
    for (int i = 0; i < 4; i++) o4[i] = r.GetFloat();
}
X float orc_GetSpecularLobeTanHalfAngle(float r, float p) { return ImportanceSampling::GetSpecularLobeTanHalfAngle(r, p); }
X float orc_GetSpecularDominantFactorG2(float NoV, float r) { return ImportanceSampling::GetSpecularDominantFactorG2(NoV, r); }
