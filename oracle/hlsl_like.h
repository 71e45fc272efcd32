// TEST INFRASTRUCTURE — CPU oracle. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may build, link or call anything under oracle/. Never part of the product path.
//
// Minimal HLSL-flavoured scalar vector types so the pass restatements in this directory can follow the
// reference shaders (External/NRD/Shaders/*.hlsl*) line by line. Plain fp32, no intrinsics, no SIMD.
// PINNED: bit-identical, dispatch by dispatch, to the reference's own shaders compiled as C++
// (oracle/_ref/libnrd_refshaders.so, tests/test_oracle_vs_reference_shaders.py).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace orc {

struct float2 {
    float x, y;
    float2() : x(0), y(0) {}
    float2(float a) : x(a), y(a) {}
    float2(float a, float b) : x(a), y(b) {}
};
struct float3 {
    float x, y, z;
    float3() : x(0), y(0), z(0) {}
    float3(float a) : x(a), y(a), z(a) {}
    float3(float a, float b, float c) : x(a), y(b), z(c) {}
    float3(float2 a, float c) : x(a.x), y(a.y), z(c) {}
    float2 xy() const { return float2(x, y); }
};
struct float4 {
    float x, y, z, w;
    float4() : x(0), y(0), z(0), w(0) {}
    float4(float a) : x(a), y(a), z(a), w(a) {}
    float4(float a, float b, float c, float d) : x(a), y(b), z(c), w(d) {}
    float4(float3 a, float d) : x(a.x), y(a.y), z(a.z), w(d) {}
    float4(float2 a, float2 b) : x(a.x), y(a.y), z(b.x), w(b.y) {}
    float3 xyz() const { return float3(x, y, z); }
    float2 xy() const { return float2(x, y); }
    float2 zw() const { return float2(z, w); }
    float& operator[](int i) { return (&x)[i]; }
    float operator[](int i) const { return (&x)[i]; }
};
struct int2 {
    int x, y;
    int2() : x(0), y(0) {}
    int2(int a, int b) : x(a), y(b) {}
};

#define ORC_OP2(T, op)                                                                     \
    inline T operator op(T a, T b) { return T(a.x op b.x, a.y op b.y); }                   \
    inline T operator op(T a, float b) { return T(a.x op b, a.y op b); }                   \
    inline T operator op(float a, T b) { return T(a op b.x, a op b.y); }
#define ORC_OP3(T, op)                                                                     \
    inline T operator op(T a, T b) { return T(a.x op b.x, a.y op b.y, a.z op b.z); }       \
    inline T operator op(T a, float b) { return T(a.x op b, a.y op b, a.z op b); }         \
    inline T operator op(float a, T b) { return T(a op b.x, a op b.y, a op b.z); }
#define ORC_OP4(T, op)                                                                               \
    inline T operator op(T a, T b) { return T(a.x op b.x, a.y op b.y, a.z op b.z, a.w op b.w); }     \
    inline T operator op(T a, float b) { return T(a.x op b, a.y op b, a.z op b, a.w op b); }         \
    inline T operator op(float a, T b) { return T(a op b.x, a op b.y, a op b.z, a op b.w); }
ORC_OP2(float2, +) ORC_OP2(float2, -) ORC_OP2(float2, *) ORC_OP2(float2, /)
ORC_OP3(float3, +) ORC_OP3(float3, -) ORC_OP3(float3, *) ORC_OP3(float3, /)
ORC_OP4(float4, +) ORC_OP4(float4, -) ORC_OP4(float4, *) ORC_OP4(float4, /)
inline float2 operator-(float2 a) { return float2(-a.x, -a.y); }
inline float3 operator-(float3 a) { return float3(-a.x, -a.y, -a.z); }
inline float4 operator-(float4 a) { return float4(-a.x, -a.y, -a.z, -a.w); }
inline float2& operator+=(float2& a, float2 b) { a = a + b; return a; }
inline float3& operator+=(float3& a, float3 b) { a = a + b; return a; }
inline float4& operator+=(float4& a, float4 b) { a = a + b; return a; }
inline float2& operator*=(float2& a, float2 b) { a = a * b; return a; }
inline float3& operator*=(float3& a, float3 b) { a = a * b; return a; }
inline float4& operator*=(float4& a, float4 b) { a = a * b; return a; }
inline float2& operator*=(float2& a, float b) { a = a * b; return a; }
inline float3& operator*=(float3& a, float b) { a = a * b; return a; }
inline float4& operator*=(float4& a, float b) { a = a * b; return a; }
inline float2& operator/=(float2& a, float b) { a = a / b; return a; }
inline float3& operator/=(float3& a, float b) { a = a / b; return a; }
inline float2& operator/=(float2& a, float2 b) { a = a / b; return a; }
inline float4& operator-=(float4& a, float b) { a = a - b; return a; }
inline float3& operator-=(float3& a, float3 b) { a = a - b; return a; }

// scalar intrinsics
inline float saturate(float x) { return x != x ? 0.0f : (x < 0.0f ? 0.0f : (x > 1.0f ? 1.0f : x)); }  // HLSL saturate(NaN) = 0
inline float lerp(float a, float b, float t) { return a + (b - a) * t; }
inline float step(float edge, float x) { return x >= edge ? 1.0f : 0.0f; }
inline float frac(float x) { return x - std::floor(x); }
inline float min(float a, float b) { return a < b ? a : b; }  // NB: not NaN-propagating, like HLSL min/max pick the non-NaN operand
inline float max(float a, float b) { return a > b ? a : b; }
inline float clamp(float x, float lo, float hi) { return min(max(x, lo), hi); }
inline float rcp(float x) { return 1.0f / x; }
inline float rsqrt(float x) { return 1.0f / std::sqrt(x); }
inline float hlsl_round(float x) { return std::nearbyint(x); }  // round-half-to-even, as DXIL Round_ne
inline float sign_fast(float x) { return step(0.0f, x) * 2.0f - 1.0f; }  // Math::Sign, ML_SIGN_FAST (ml.hlsli:166)

#define ORC_MAP2(name) inline float2 name(float2 a) { return float2(name(a.x), name(a.y)); }
#define ORC_MAP3(name) inline float3 name(float3 a) { return float3(name(a.x), name(a.y), name(a.z)); }
#define ORC_MAP4(name) inline float4 name(float4 a) { return float4(name(a.x), name(a.y), name(a.z), name(a.w)); }
ORC_MAP2(saturate) ORC_MAP3(saturate) ORC_MAP4(saturate)
ORC_MAP2(frac)
inline float2 abs(float2 a) { return float2(std::fabs(a.x), std::fabs(a.y)); }
inline float3 abs(float3 a) { return float3(std::fabs(a.x), std::fabs(a.y), std::fabs(a.z)); }
inline float4 abs(float4 a) { return float4(std::fabs(a.x), std::fabs(a.y), std::fabs(a.z), std::fabs(a.w)); }
inline float2 floor(float2 a) { return float2(std::floor(a.x), std::floor(a.y)); }
inline float2 min(float2 a, float2 b) { return float2(min(a.x, b.x), min(a.y, b.y)); }
inline float2 max(float2 a, float2 b) { return float2(max(a.x, b.x), max(a.y, b.y)); }
inline float3 min(float3 a, float3 b) { return float3(min(a.x, b.x), min(a.y, b.y), min(a.z, b.z)); }
inline float3 max(float3 a, float3 b) { return float3(max(a.x, b.x), max(a.y, b.y), max(a.z, b.z)); }
inline float4 min(float4 a, float4 b) { return float4(min(a.x, b.x), min(a.y, b.y), min(a.z, b.z), min(a.w, b.w)); }
inline float4 max(float4 a, float4 b) { return float4(max(a.x, b.x), max(a.y, b.y), max(a.z, b.z), max(a.w, b.w)); }
inline float2 lerp(float2 a, float2 b, float t) { return a + (b - a) * t; }
inline float2 lerp(float2 a, float2 b, float2 t) { return a + (b - a) * t; }
inline float3 lerp(float3 a, float3 b, float t) { return a + (b - a) * t; }
inline float4 lerp(float4 a, float4 b, float t) { return a + (b - a) * t; }
inline float4 lerp(float4 a, float4 b, float4 t) { return a + (b - a) * t; }
inline float2 step(float2 e, float2 x) { return float2(step(e.x, x.x), step(e.y, x.y)); }
inline float3 step(float3 e, float3 x) { return float3(step(e.x, x.x), step(e.y, x.y), step(e.z, x.z)); }
inline float4 step(float4 e, float4 x) { return float4(step(e.x, x.x), step(e.y, x.y), step(e.z, x.z), step(e.w, x.w)); }
inline float dot(float2 a, float2 b) { return a.x * b.x + a.y * b.y; }
inline float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float dot(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
inline float length(float2 a) { return std::sqrt(dot(a, a)); }
inline float length(float3 a) { return std::sqrt(dot(a, a)); }
inline float3 normalize(float3 a) { return a * rsqrt(dot(a, a)); }   // DXC lowers normalize to v * rsqrt( dot( v, v ) )
inline float3 cross(float3 a, float3 b) { return float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
inline float3 reflect(float3 i, float3 n) { return i - 2.0f * n * dot(i, n); }

inline uint32_t asuint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
inline float asfloat(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

// IEEE binary16 <-> binary32, round-to-nearest-even on narrowing (what an RGBA16F store does)
inline uint16_t f32tof16(float f) {
    uint32_t x = asuint(f);
    uint32_t sign = (x >> 16) & 0x8000u;
    uint32_t ax = x & 0x7FFFFFFFu;
    if (ax >= 0x7F800000u) return (uint16_t)(sign | (ax > 0x7F800000u ? 0x7E00u : 0x7C00u));  // NaN / Inf
    if (ax >= 0x477FF000u) return (uint16_t)(sign | 0x7C00u);                                    // rounds to Inf (>= 65520)
    if (ax < 0x33000001u) return (uint16_t)sign;                                                 // rounds to zero (<= 2^-25)
    int32_t exp = (int32_t)(ax >> 23) - 127;
    uint32_t man = (ax & 0x7FFFFFu) | 0x800000u;
    uint32_t shift, half;
    if (exp < -14) {  // subnormal half
        shift = (uint32_t)(13 + (-14 - exp));
        half = 0;
    } else {
        shift = 13;
        half = (uint32_t)(exp + 15) << 10;
    }
    uint32_t q = man >> shift;
    uint32_t rem = man & ((1u << shift) - 1u);
    uint32_t halfway = 1u << (shift - 1);
    if (rem > halfway || (rem == halfway && (q & 1u))) q++;
    if (exp >= -14) q -= 0x400u;  // drop the implicit bit (a carry out of the mantissa bumps the exponent correctly)
    return (uint16_t)(sign | (half + q));
}
inline float f16tof32(uint32_t h) {
    h &= 0xFFFFu;
    uint32_t sign = (h & 0x8000u) << 16, exp = (h >> 10) & 0x1Fu, man = h & 0x3FFu;
    if (exp == 0) {
        if (man == 0) return asfloat(sign);
        float v = std::ldexp((float)man, -24);
        return sign ? -v : v;
    }
    if (exp == 31) return asfloat(sign | 0x7F800000u | (man << 13));
    return asfloat(sign | ((exp + 112u) << 23) | (man << 13));
}

// Column-major 4x4 as it sits in the constant buffers (element(r, c) = m[c * 4 + r]); mul(M, v) with v a column
struct float4x4 {
    float m[16];
};
inline float4 mul(const float4x4& M, float4 v) {
    return float4(M.m[0] * v.x + M.m[4] * v.y + M.m[8] * v.z + M.m[12] * v.w, M.m[1] * v.x + M.m[5] * v.y + M.m[9] * v.z + M.m[13] * v.w,
                  M.m[2] * v.x + M.m[6] * v.y + M.m[10] * v.z + M.m[14] * v.w, M.m[3] * v.x + M.m[7] * v.y + M.m[11] * v.z + M.m[15] * v.w);
}
inline float3 mul3x3(const float4x4& M, float3 v) {  // (float3x3)M * v
    return float3(M.m[0] * v.x + M.m[4] * v.y + M.m[8] * v.z, M.m[1] * v.x + M.m[5] * v.y + M.m[9] * v.z, M.m[2] * v.x + M.m[6] * v.y + M.m[10] * v.z);
}
inline float3 mul3x3T(const float4x4& M, float3 v) {  // transpose((float3x3)M) * v
    return float3(M.m[0] * v.x + M.m[1] * v.y + M.m[2] * v.z, M.m[4] * v.x + M.m[5] * v.y + M.m[6] * v.z, M.m[8] * v.x + M.m[9] * v.y + M.m[10] * v.z);
}

}  // namespace orc
