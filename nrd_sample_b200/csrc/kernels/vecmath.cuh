// Device-side vector arithmetic on CUDA's built-in float2/float3/float4 plus the scalar helpers the
// denoiser kernels use everywhere. fp32 throughout; no fast-math substitutions here — kernels opt into
// approximate intrinsics explicitly where the parity budget allows.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>

#define NRD_DEV __device__ __forceinline__

namespace nrdk {

// ---- constructors -------------------------------------------------------------------------------
NRD_DEV float2 f2(float a) { return make_float2(a, a); }
NRD_DEV float2 f2(float a, float b) { return make_float2(a, b); }
NRD_DEV float3 f3(float a) { return make_float3(a, a, a); }
NRD_DEV float3 f3(float a, float b, float c) { return make_float3(a, b, c); }
NRD_DEV float4 f4(float a) { return make_float4(a, a, a, a); }
NRD_DEV float4 f4(float a, float b, float c, float d) { return make_float4(a, b, c, d); }
NRD_DEV float4 f4(float3 a, float d) { return make_float4(a.x, a.y, a.z, d); }
NRD_DEV float3 xyz(float4 a) { return make_float3(a.x, a.y, a.z); }
NRD_DEV float2 xy(float3 a) { return make_float2(a.x, a.y); }
NRD_DEV float2 xy(float4 a) { return make_float2(a.x, a.y); }

// ---- operators ----------------------------------------------------------------------------------
#define NRDK_OPS(op)                                                                                                   \
    NRD_DEV float2 operator op(float2 a, float2 b) { return make_float2(a.x op b.x, a.y op b.y); }                     \
    NRD_DEV float2 operator op(float2 a, float b) { return make_float2(a.x op b, a.y op b); }                          \
    NRD_DEV float2 operator op(float a, float2 b) { return make_float2(a op b.x, a op b.y); }                          \
    NRD_DEV float3 operator op(float3 a, float3 b) { return make_float3(a.x op b.x, a.y op b.y, a.z op b.z); }         \
    NRD_DEV float3 operator op(float3 a, float b) { return make_float3(a.x op b, a.y op b, a.z op b); }                \
    NRD_DEV float3 operator op(float a, float3 b) { return make_float3(a op b.x, a op b.y, a op b.z); }                \
    NRD_DEV float4 operator op(float4 a, float4 b) { return make_float4(a.x op b.x, a.y op b.y, a.z op b.z, a.w op b.w); } \
    NRD_DEV float4 operator op(float4 a, float b) { return make_float4(a.x op b, a.y op b, a.z op b, a.w op b); }      \
    NRD_DEV float4 operator op(float a, float4 b) { return make_float4(a op b.x, a op b.y, a op b.z, a op b.w); }
NRDK_OPS(+)
NRDK_OPS(-)
NRDK_OPS(*)
NRDK_OPS(/)
#undef NRDK_OPS
NRD_DEV float2 operator-(float2 a) { return make_float2(-a.x, -a.y); }
NRD_DEV float3 operator-(float3 a) { return make_float3(-a.x, -a.y, -a.z); }
NRD_DEV float4 operator-(float4 a) { return make_float4(-a.x, -a.y, -a.z, -a.w); }
template <class T, class U> NRD_DEV T& operator+=(T& a, U b) { a = a + b; return a; }
template <class T, class U> NRD_DEV T& operator-=(T& a, U b) { a = a - b; return a; }
template <class T, class U> NRD_DEV T& operator*=(T& a, U b) { a = a * b; return a; }
template <class T, class U> NRD_DEV T& operator/=(T& a, U b) { a = a / b; return a; }

// ---- scalar helpers -----------------------------------------------------------------------------
NRD_DEV float saturate(float x) { return __saturatef(x); }  // NaN -> 0, like HLSL
NRD_DEV float lerp(float a, float b, float t) { return a + (b - a) * t; }
NRD_DEV float step(float edge, float x) { return x >= edge ? 1.0f : 0.0f; }
NRD_DEV float frac(float x) { return x - floorf(x); }
NRD_DEV float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
NRD_DEV float signFast(float x) { return x >= 0.0f ? 1.0f : -1.0f; }  // step(0, x) * 2 - 1
NRD_DEV float roundNe(float x) { return rintf(x); }
NRD_DEV int clampi(int v, int lo, int hi) { return max(lo, min(v, hi)); }

NRD_DEV float2 saturate(float2 a) { return make_float2(saturate(a.x), saturate(a.y)); }
NRD_DEV float4 saturate(float4 a) { return make_float4(saturate(a.x), saturate(a.y), saturate(a.z), saturate(a.w)); }
NRD_DEV float2 fabs2(float2 a) { return make_float2(fabsf(a.x), fabsf(a.y)); }
NRD_DEV float3 fabs3(float3 a) { return make_float3(fabsf(a.x), fabsf(a.y), fabsf(a.z)); }
NRD_DEV float4 fabs4(float4 a) { return make_float4(fabsf(a.x), fabsf(a.y), fabsf(a.z), fabsf(a.w)); }
NRD_DEV float2 floor2(float2 a) { return make_float2(floorf(a.x), floorf(a.y)); }
NRD_DEV float2 frac2(float2 a) { return make_float2(frac(a.x), frac(a.y)); }
NRD_DEV float2 min2(float2 a, float2 b) { return make_float2(fminf(a.x, b.x), fminf(a.y, b.y)); }
NRD_DEV float2 max2(float2 a, float2 b) { return make_float2(fmaxf(a.x, b.x), fmaxf(a.y, b.y)); }
NRD_DEV float3 max3(float3 a, float3 b) { return make_float3(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)); }
NRD_DEV float2 lerp(float2 a, float2 b, float t) { return a + (b - a) * t; }
NRD_DEV float2 lerp(float2 a, float2 b, float2 t) { return a + (b - a) * t; }
NRD_DEV float3 lerp(float3 a, float3 b, float t) { return a + (b - a) * t; }
NRD_DEV float4 lerp(float4 a, float4 b, float t) { return a + (b - a) * t; }
NRD_DEV float4 lerp(float4 a, float4 b, float4 t) { return a + (b - a) * t; }
NRD_DEV float dot(float2 a, float2 b) { return a.x * b.x + a.y * b.y; }
NRD_DEV float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
NRD_DEV float dot(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
NRD_DEV float sum4(float4 a) { return a.x + a.y + a.z + a.w; }
NRD_DEV float length(float2 a) { return sqrtf(dot(a, a)); }
NRD_DEV float length(float3 a) { return sqrtf(dot(a, a)); }
NRD_DEV float3 normalize(float3 a) { return a / length(a); }
NRD_DEV float3 cross(float3 a, float3 b) { return make_float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
NRD_DEV float3 reflect(float3 i, float3 n) { return i - 2.0f * n * dot(i, n); }

// saturate( 1 - |x| ) as ONE FADD.SAT with the |.| and negate modifiers folded in. Written in C, nvcc emits a separate FADD for the absolute value ( the
// .ftz flavour of abs cannot be a source modifier ) and another for the subtraction before the saturate: 3 instructions for the same bits.
#ifdef __CUDA_ARCH__
NRD_DEV float satOneMinusAbs(float x) {
    float d;
    asm("{\n\t.reg .f32 t;\n\tabs.f32 t, %1;\n\tsub.sat.ftz.f32 %0, 0f3F800000, t;\n\t}" : "=f"(d) : "f"(x));
    return d;
}
#else
NRD_DEV float satOneMinusAbs(float x) { return fminf(fmaxf(1.0f - fabsf(x), 0.0f), 1.0f); }
#endif

NRD_DEV float comp(float4 v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w)); }

}  // namespace nrdk
