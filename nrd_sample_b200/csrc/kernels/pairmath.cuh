// Packed fp32x2 arithmetic for sm_100a. Blackwell issues FFMA2 / FMUL2 / FADD2 — one warp instruction, two fp32 lanes per
// thread — and the denoiser's gather loops are issue-bound on scalar FADD/FMUL/FFMA (profiles/README.md), so the spatial
// kernels evaluate TWO Poisson taps per thread in lock step: every per-tap quantity is a P2 holding (tap a, tap b).
//
// Packed instructions take no |x| / -x / .SAT operand modifiers, so a step that needs one is issued as two scalar
// instructions WITH the modifier folded in (e.g. sat(1 - |x|) is one FADD.SAT per lane) rather than packed + fix-up;
// subtraction is fma2(y, -1, x). Loads, conversions, MUFU, min/max and compares are scalar per lane as well.
#pragma once
#include "vecmath.cuh"

namespace nrdk {

struct P2 {
    float2 v;
    NRD_DEV P2() {}
    NRD_DEV explicit P2(float a) : v(make_float2(a, a)) {}
    NRD_DEV P2(float a, float b) : v(make_float2(a, b)) {}
    NRD_DEV explicit P2(float2 a) : v(a) {}
    NRD_DEV float a() const { return v.x; }
    NRD_DEV float b() const { return v.y; }
};

NRD_DEV P2 toP2(P2 x) { return x; }
NRD_DEV P2 toP2(float x) { return P2(x); }

// x * y + z with any mix of pairs and (broadcast) scalars
template <class A, class B, class C> NRD_DEV P2 fma2(A x, B y, C z) { return P2(__ffma2_rn(toP2(x).v, toP2(y).v, toP2(z).v)); }
NRD_DEV P2 operator+(P2 x, P2 y) { return P2(__fadd2_rn(x.v, y.v)); }
NRD_DEV P2 operator*(P2 x, P2 y) { return P2(__fmul2_rn(x.v, y.v)); }
NRD_DEV P2 operator-(P2 x, P2 y) { return fma2(y, -1.0f, x); }
NRD_DEV P2 operator+(P2 x, float y) { return x + P2(y); }
NRD_DEV P2 operator+(float x, P2 y) { return P2(x) + y; }
NRD_DEV P2 operator-(P2 x, float y) { return x + P2(-y); }
NRD_DEV P2 operator-(float x, P2 y) { return fma2(y, -1.0f, x); }
NRD_DEV P2 operator*(P2 x, float y) { return x * P2(y); }
NRD_DEV P2 operator*(float x, P2 y) { return P2(x) * y; }

// x * y rounded to fp32 and only THEN + z ( two roundings, as an IEEE engine without contraction does ). Neither the __fmul2_rn / __fadd2_rn intrinsics nor
// explicit mul.rn.f32x2 / add.rn.f32x2 PTX keep the packed pair apart under -use_fast_math ( measured: one FFMA2 either way ); the scalar __fadd_rn is
// documented never to fuse, so the addition is issued per lane.
NRD_DEV P2 mulThenAdd2(P2 x, P2 y, float z) {
    const P2 p = x * y;
    return P2(__fadd_rn(p.v.x, z), __fadd_rn(p.v.y, z));
}

// scalar-per-lane steps, each written so that one instruction per lane suffices
NRD_DEV float satOneMinus(float x) {  // saturate( 1 - x ) as ONE FADD.SAT (nvcc otherwise emits 1 - x and the saturate separately)
    float d;
    asm("sub.sat.ftz.f32 %0, 0f3F800000, %1;" : "=f"(d) : "f"(x));
    return d;
}
NRD_DEV P2 sat2(P2 x) { return P2(__saturatef(x.v.x), __saturatef(x.v.y)); }                    // FADD.SAT
NRD_DEV P2 satOneMinus2(P2 x) { return P2(satOneMinus(x.v.x), satOneMinus(x.v.y)); }            // FADD.SAT 1, -x
// ( satOneMinusAbs: vecmath.cuh — 2 instructions per pair; 1 - min( |x|, 1 ) as two FMNMX and a packed subtraction is 3 and gives the same bits )
NRD_DEV P2 oneMinusAbsSat2(P2 x) { return P2(satOneMinusAbs(x.v.x), satOneMinusAbs(x.v.y)); }
NRD_DEV P2 absMul2(P2 x, float k) { return P2(fabsf(x.v.x) * k, fabsf(x.v.y) * k); }                                         // FMUL |x|, k
NRD_DEV P2 mulSat2(P2 x, P2 y) { return P2(__saturatef(x.v.x * y.v.x), __saturatef(x.v.y * y.v.y)); }                        // FMUL.SAT
NRD_DEV P2 abs2(P2 x) { return P2(fabsf(x.v.x), fabsf(x.v.y)); }
NRD_DEV P2 min2(P2 x, float y) { return P2(fminf(x.v.x, y), fminf(x.v.y, y)); }
NRD_DEV P2 max2(P2 x, float y) { return P2(fmaxf(x.v.x, y), fmaxf(x.v.y, y)); }
NRD_DEV P2 floor2(P2 x) { return P2(floorf(x.v.x), floorf(x.v.y)); }
NRD_DEV P2 rcp2(P2 x) { return P2(1.0f / x.v.x, 1.0f / x.v.y); }
NRD_DEV P2 rsqrt2(P2 x) { return P2(rsqrtf(x.v.x), rsqrtf(x.v.y)); }
NRD_DEV P2 sqrt2(P2 x) { return P2(sqrtf(x.v.x), sqrtf(x.v.y)); }
NRD_DEV P2 sel2(bool ca, bool cb, P2 t, P2 f) { return P2(ca ? t.v.x : f.v.x, cb ? t.v.y : f.v.y); }

}  // namespace nrdk
