// TEST INFRASTRUCTURE — a small HLSL-on-CPU runtime.
//
// Purpose: compile the reference's own compute shaders (External/NRD/Shaders/*.cs.hlsl + MathLib's ml.hlsli), from where they lie
// under /root/reference, as C++ into oracle/_ref/libnrd_refshaders.so, so that the hand-written oracle (oracle/*_passes.cpp) can be
// pinned against the reference's real arithmetic. The shader text is piped through the C preprocessor and oracle/ref_shim/hlsl2cpp.py
// straight into g++ (nothing of it is stored in this repository); this header supplies what HLSL has and C++ lacks:
//   * vectors with swizzles, matrices, implicit scalar<->vector promotion, the intrinsics the shaders use
//   * Texture2D / RWTexture2D / SamplerState over pitch-linear host memory in the nrd::Format encodings
//   * cbuffer packing, resource binding in register order
//   * a thread group executed as fibers so that GroupMemoryBarrierWithGroupSync() and SM6.0 quad reads work
// Nothing here knows anything about NRD's algorithms.
#pragma once
#include <ucontext.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <type_traits>
#include <vector>

namespace hlsl {
typedef unsigned int uint;

template <class T, int N> struct vec;
template <class T, int N, int... I> struct Swz;

// ------------------------------------------------------------------------------------------------ traits
template <class A> struct Tr { static constexpr bool num = false, hl = false; static constexpr int N = 0; typedef void E; };
#define HLSL_SCALAR_TRAIT(T, ET) \
    template <> struct Tr<T> { static constexpr bool num = true, hl = false; static constexpr int N = 1; typedef ET E; static E get(const T& a, int) { return (E)a; } };
HLSL_SCALAR_TRAIT(float, float)
HLSL_SCALAR_TRAIT(double, float)
HLSL_SCALAR_TRAIT(int, int)
HLSL_SCALAR_TRAIT(uint, uint)
HLSL_SCALAR_TRAIT(bool, bool)
HLSL_SCALAR_TRAIT(long, int)
HLSL_SCALAR_TRAIT(unsigned long, uint)
// what a Texture2D<float> / Texture2D<uint> fetch returns: a scalar that also answers to ".x" / ".r" as in HLSL
template <class T> struct Scalar1 { union { T x; T r; }; explicit Scalar1(T v) : x(v) {} operator T() const { return x; } };
template <class T> struct Tr<Scalar1<T>> { static constexpr bool num = true, hl = false; static constexpr int N = 1; typedef T E; static T get(const Scalar1<T>& a, int) { return a.x; } };
template <class T, int M> struct Tr<vec<T, M>> {
    static constexpr bool num = true, hl = true; static constexpr int N = M; typedef T E;
    static T get(const vec<T, M>& a, int i) { return a.d[i]; }
};
template <class T, int M, int... I> struct Tr<Swz<T, M, I...>> {
    static constexpr bool num = true, hl = true; static constexpr int N = sizeof...(I); typedef T E;
    static T get(const Swz<T, M, I...>& a, int i) { constexpr int idx[] = {I...}; return a.d[idx[i]]; }
};

template <class EA, class EB> struct Prom2 { typedef int E; };
template <class EB> struct Prom2<float, EB> { typedef float E; };
template <> struct Prom2<int, float> { typedef float E; };
template <> struct Prom2<uint, float> { typedef float E; };
template <> struct Prom2<bool, float> { typedef float E; };
template <> struct Prom2<uint, uint> { typedef uint E; };
template <> struct Prom2<uint, int> { typedef uint E; };
template <> struct Prom2<uint, bool> { typedef uint E; };
template <> struct Prom2<int, uint> { typedef uint E; };
template <> struct Prom2<bool, uint> { typedef uint E; };
template <class A, class B> using Prom = typename Prom2<typename Tr<A>::E, typename Tr<B>::E>::E;
template <class A, class B> constexpr int maxN() {
    static_assert(Tr<A>::N == Tr<B>::N || Tr<A>::N == 1 || Tr<B>::N == 1, "vector sizes do not match");
    return Tr<A>::N > Tr<B>::N ? Tr<A>::N : Tr<B>::N;
}
template <class E, int N> using RT = std::conditional_t<N == 1, E, vec<E, N>>;
template <class E, int N, class F> inline RT<E, N> build(F f) {
    if constexpr (N == 1) return (E)f(0);
    else { vec<E, N> r; for (int i = 0; i < N; i++) r.d[i] = (E)f(i); return r; }
}
#define HLSL_NUM(A) class = std::enable_if_t<Tr<A>::num>
#define HLSL_NUM2(A, B) class = std::enable_if_t<Tr<A>::num && Tr<B>::num>
#define HLSL_NUM3(A, B, C) class = std::enable_if_t<Tr<A>::num && Tr<B>::num && Tr<C>::num>
#define HLSL_ANYVEC2(A, B) class = std::enable_if_t<Tr<A>::num && Tr<B>::num && (Tr<A>::hl || Tr<B>::hl)>

// ------------------------------------------------------------------------------------------------ swizzles
template <class T, int N, int... I> struct Swz {
    T d[N];
    static constexpr int K = sizeof...(I);
    vec<T, K> v() const { vec<T, K> r; constexpr int idx[] = {I...}; for (int i = 0; i < K; i++) r.d[i] = d[idx[i]]; return r; }
    template <class A, HLSL_NUM(A)> Swz& operator=(const A& a) {
        static_assert(Tr<A>::N == 1 || Tr<A>::N == K, "swizzle assignment size");
        constexpr int idx[] = {I...}; T tmp[K];
        for (int i = 0; i < K; i++) tmp[i] = (T)Tr<A>::get(a, i);
        for (int i = 0; i < K; i++) d[idx[i]] = tmp[i];
        return *this;
    }
    Swz& operator=(const Swz& o) { return *this = o.v(); }
    template <class A> Swz& operator+=(const A& a) { return *this = v() + a; }
    template <class A> Swz& operator-=(const A& a) { return *this = v() - a; }
    template <class A> Swz& operator*=(const A& a) { return *this = v() * a; }
    template <class A> Swz& operator/=(const A& a) { return *this = v() / a; }
    template <class A> Swz& operator%=(const A& a) { return *this = v() % a; }
    template <class A> Swz& operator&=(const A& a) { return *this = v() & a; }
    template <class A> Swz& operator|=(const A& a) { return *this = v() | a; }
    template <class A> Swz& operator^=(const A& a) { return *this = v() ^ a; }
    template <class A> Swz& operator<<=(const A& a) { return *this = v() << a; }
    template <class A> Swz& operator>>=(const A& a) { return *this = v() >> a; }
    T operator[](int i) const { constexpr int idx[] = {I...}; return d[idx[i]]; }
};
#include "hlsl_swizzles.inc"

// ------------------------------------------------------------------------------------------------ vectors
template <int N, class... A> constexpr int totalN() { return (0 + ... + Tr<A>::N); }
template <class T, int N, class A> inline void flattenInto(T* d, int& k, const A& a) { for (int i = 0; i < Tr<A>::N; i++) d[k++] = (T)Tr<A>::get(a, i); }

#define HLSL_VEC_COMMON(NN) \
    vec() { for (int i = 0; i < NN; i++) d[i] = T(0); } \
    vec(const vec& o) { for (int i = 0; i < NN; i++) d[i] = o.d[i]; } \
    vec& operator=(const vec& o) { for (int i = 0; i < NN; i++) d[i] = o.d[i]; return *this; } \
    /* scalar broadcast / same-size conversion: implicit */ \
    template <class A, std::enable_if_t<Tr<A>::num && (Tr<A>::N == 1 || Tr<A>::N == NN), int> = 0> vec(const A& a) { for (int i = 0; i < NN; i++) d[i] = (T)Tr<A>::get(a, i); } \
    /* truncation: explicit, ( float3 )v4 */ \
    template <class A, std::enable_if_t<Tr<A>::num && (Tr<A>::N > NN), long> = 0> explicit vec(const A& a) { for (int i = 0; i < NN; i++) d[i] = (T)Tr<A>::get(a, i); } \
    /* float4( v3, s ), float4( v2, v2 ), float3( a, b, c ) ... */ \
    template <class A0, class A1, class... AR, std::enable_if_t<(Tr<A0>::num && Tr<A1>::num && (... && Tr<AR>::num)) && (Tr<A0>::N + Tr<A1>::N + (0 + ... + Tr<AR>::N) == NN), int> = 0> \
    vec(const A0& a0, const A1& a1, const AR&... ar) { int k = 0; flattenInto<T, NN>(d, k, a0); flattenInto<T, NN>(d, k, a1); (flattenInto<T, NN>(d, k, ar), ...); } \
    T& operator[](int i) { return d[i]; } \
    const T& operator[](int i) const { return d[i]; } \
    template <class A> vec& operator+=(const A& a) { return *this = vec(*this + a); } \
    template <class A> vec& operator-=(const A& a) { return *this = vec(*this - a); } \
    template <class A> vec& operator*=(const A& a) { return *this = vec(*this * a); } \
    template <class A> vec& operator/=(const A& a) { return *this = vec(*this / a); } \
    template <class A> vec& operator%=(const A& a) { return *this = vec(*this % a); } \
    template <class A> vec& operator&=(const A& a) { return *this = vec(*this & a); } \
    template <class A> vec& operator|=(const A& a) { return *this = vec(*this | a); } \
    template <class A> vec& operator^=(const A& a) { return *this = vec(*this ^ a); } \
    template <class A> vec& operator<<=(const A& a) { return *this = vec(*this << a); } \
    template <class A> vec& operator>>=(const A& a) { return *this = vec(*this >> a); }

template <class T> struct vec<T, 2> {
    union { T d[2]; struct { T x, y; }; struct { T r, g; }; HLSL_SWIZZLES_2(T) };
    HLSL_VEC_COMMON(2)
};
template <class T> struct vec<T, 3> {
    union { T d[3]; struct { T x, y, z; }; struct { T r, g, b; }; HLSL_SWIZZLES_3(T) };
    HLSL_VEC_COMMON(3)
};
template <class T> struct vec<T, 4> {
    union { T d[4]; struct { T x, y, z, w; }; struct { T r, g, b, a; }; HLSL_SWIZZLES_4(T) };
    HLSL_VEC_COMMON(4)
};
typedef vec<float, 2> float2; typedef vec<float, 3> float3; typedef vec<float, 4> float4;
typedef vec<int, 2> int2; typedef vec<int, 3> int3; typedef vec<int, 4> int4;
typedef vec<uint, 2> uint2; typedef vec<uint, 3> uint3; typedef vec<uint, 4> uint4;
typedef vec<bool, 2> bool2; typedef vec<bool, 3> bool3; typedef vec<bool, 4> bool4;

// ------------------------------------------------------------------------------------------------ operators
#define HLSL_ARITH(op) \
    template <class A, class B, HLSL_ANYVEC2(A, B)> inline auto operator op(const A& a, const B& b) { \
        typedef Prom<A, B> E; constexpr int N = maxN<A, B>(); \
        return build<E, N>([&](int i) { return (E)((E)Tr<A>::get(a, i) op (E)Tr<B>::get(b, i)); }); }
HLSL_ARITH(+) HLSL_ARITH(-) HLSL_ARITH(*) HLSL_ARITH(&) HLSL_ARITH(|) HLSL_ARITH(^)
template <class A, class B, HLSL_ANYVEC2(A, B)> inline auto operator/(const A& a, const B& b) {
    typedef Prom<A, B> E; constexpr int N = maxN<A, B>();
    return build<E, N>([&](int i) { E y = (E)Tr<B>::get(b, i); if constexpr (!std::is_floating_point_v<E>) { if (y == 0) return (E)0; } return (E)((E)Tr<A>::get(a, i) / y); }); }
template <class A, class B, HLSL_ANYVEC2(A, B)> inline auto operator%(const A& a, const B& b) {
    typedef Prom<A, B> E; constexpr int N = maxN<A, B>();
    return build<E, N>([&](int i) { E x = (E)Tr<A>::get(a, i), y = (E)Tr<B>::get(b, i);
        if constexpr (std::is_floating_point_v<E>) return (E)std::fmod(x, y); else return (E)(y == 0 ? 0 : x % y); }); }
template <class A, class B, HLSL_ANYVEC2(A, B)> inline auto operator<<(const A& a, const B& b) {
    typedef typename Tr<A>::E E; constexpr int N = maxN<A, B>();
    return build<E, N>([&](int i) { return (E)(Tr<A>::get(a, i) << ((uint)Tr<B>::get(b, i) & 31u)); }); }
template <class A, class B, HLSL_ANYVEC2(A, B)> inline auto operator>>(const A& a, const B& b) {
    typedef typename Tr<A>::E E; constexpr int N = maxN<A, B>();
    return build<E, N>([&](int i) { return (E)(Tr<A>::get(a, i) >> ((uint)Tr<B>::get(b, i) & 31u)); }); }
#define HLSL_CMP(op) \
    template <class A, class B, HLSL_ANYVEC2(A, B)> inline auto operator op(const A& a, const B& b) { \
        typedef Prom<A, B> E; constexpr int N = maxN<A, B>(); \
        return build<bool, N>([&](int i) { return (E)Tr<A>::get(a, i) op (E)Tr<B>::get(b, i); }); }
HLSL_CMP(<) HLSL_CMP(<=) HLSL_CMP(>) HLSL_CMP(>=) HLSL_CMP(==) HLSL_CMP(!=)
template <class A, class B, HLSL_ANYVEC2(A, B)> inline auto operator&&(const A& a, const B& b) {
    constexpr int N = maxN<A, B>(); return build<bool, N>([&](int i) { return (bool)Tr<A>::get(a, i) && (bool)Tr<B>::get(b, i); }); }
template <class A, class B, HLSL_ANYVEC2(A, B)> inline auto operator||(const A& a, const B& b) {
    constexpr int N = maxN<A, B>(); return build<bool, N>([&](int i) { return (bool)Tr<A>::get(a, i) || (bool)Tr<B>::get(b, i); }); }
template <class A, class = std::enable_if_t<Tr<A>::hl>> inline auto operator-(const A& a) {
    typedef typename Tr<A>::E E; return build<E, Tr<A>::N>([&](int i) { return (E)(-Tr<A>::get(a, i)); }); }
template <class A, class = std::enable_if_t<Tr<A>::hl>> inline auto operator+(const A& a) {
    typedef typename Tr<A>::E E; return build<E, Tr<A>::N>([&](int i) { return Tr<A>::get(a, i); }); }
template <class A, class = std::enable_if_t<Tr<A>::hl>> inline auto operator!(const A& a) {
    return build<bool, Tr<A>::N>([&](int i) { return !(bool)Tr<A>::get(a, i); }); }
template <class A, class = std::enable_if_t<Tr<A>::hl>> inline auto operator~(const A& a) {
    typedef typename Tr<A>::E E; return build<E, Tr<A>::N>([&](int i) { return (E)(~Tr<A>::get(a, i)); }); }

// ------------------------------------------------------------------------------------------------ intrinsics
#define HLSL_F1(name, expr) \
    template <class A, HLSL_NUM(A)> inline auto name(const A& a) { return build<float, Tr<A>::N>([&](int i) { float x = (float)Tr<A>::get(a, i); return (float)(expr); }); }
#define HLSL_F2(name, expr) \
    template <class A, class B, HLSL_NUM2(A, B)> inline auto name(const A& a, const B& b) { constexpr int N = maxN<A, B>(); \
        return build<float, N>([&](int i) { float x = (float)Tr<A>::get(a, i), y = (float)Tr<B>::get(b, i); return (float)(expr); }); }
#define HLSL_F3(name, expr) \
    template <class A, class B, class C, HLSL_NUM3(A, B, C)> inline auto name(const A& a, const B& b, const C& c) { constexpr int N = maxN<RT<float, maxN<A, B>()>, C>(); \
        return build<float, N>([&](int i) { float x = (float)Tr<A>::get(a, i), y = (float)Tr<B>::get(b, i), z = (float)Tr<C>::get(c, i); return (float)(expr); }); }
HLSL_F1(sqrt, std::sqrt(x)) HLSL_F1(rsqrt, 1.0f / std::sqrt(x)) HLSL_F1(rcp, 1.0f / x) HLSL_F1(floor, std::floor(x)) HLSL_F1(ceil, std::ceil(x))
HLSL_F1(frac, x - std::floor(x)) HLSL_F1(round, std::nearbyint(x)) HLSL_F1(trunc, std::trunc(x)) HLSL_F1(exp, std::exp(x)) HLSL_F1(exp2, std::exp2(x))
HLSL_F1(log, std::log(x)) HLSL_F1(log2, std::log2(x)) HLSL_F1(sin, std::sin(x)) HLSL_F1(cos, std::cos(x)) HLSL_F1(tan, std::tan(x)) HLSL_F1(asin, std::asin(x))
HLSL_F1(acos, std::acos(x)) HLSL_F1(atan, std::atan(x)) HLSL_F1(saturate, std::fmin(std::fmax(x, 0.0f), 1.0f)) HLSL_F1(radians, x * 0.017453292519943295f)
HLSL_F1(degrees, x * 57.29577951308232f)
HLSL_F2(pow, std::pow(x, y)) HLSL_F2(atan2, std::atan2(x, y)) HLSL_F2(fmod, std::fmod(x, y)) HLSL_F2(step, y >= x ? 1.0f : 0.0f)
HLSL_F3(lerp, x + (y - x) * z) HLSL_F3(mad, x * y + z)
HLSL_F3(smoothstep, [&] { float t = std::fmin(std::fmax((z - x) / (y - x), 0.0f), 1.0f); return t * t * (3.0f - 2.0f * t); }())

template <class E> inline E minE(E x, E y) { if constexpr (std::is_floating_point_v<E>) return std::fmin(x, y); else return x < y ? x : y; }
template <class E> inline E maxE(E x, E y) { if constexpr (std::is_floating_point_v<E>) return std::fmax(x, y); else return x > y ? x : y; }
template <class A, class B, HLSL_NUM2(A, B)> inline auto min(const A& a, const B& b) {
    typedef Prom<A, B> E; return build<E, maxN<A, B>()>([&](int i) { return minE<E>((E)Tr<A>::get(a, i), (E)Tr<B>::get(b, i)); }); }
template <class A, class B, HLSL_NUM2(A, B)> inline auto max(const A& a, const B& b) {
    typedef Prom<A, B> E; return build<E, maxN<A, B>()>([&](int i) { return maxE<E>((E)Tr<A>::get(a, i), (E)Tr<B>::get(b, i)); }); }
template <class A, class B, class C, HLSL_NUM3(A, B, C)> inline auto clamp(const A& a, const B& lo, const C& hi) { return min(max(a, lo), hi); }
template <class A, HLSL_NUM(A)> inline auto abs(const A& a) {
    typedef typename Tr<A>::E E; return build<E, Tr<A>::N>([&](int i) { E x = Tr<A>::get(a, i); if constexpr (std::is_floating_point_v<E>) return std::fabs(x); else return (E)(x < 0 ? -x : x); }); }
template <class A, HLSL_NUM(A)> inline auto sign(const A& a) {
    typedef typename Tr<A>::E E; return build<E, Tr<A>::N>([&](int i) { auto x = Tr<A>::get(a, i); return (E)(x > 0 ? 1 : (x < 0 ? -1 : 0)); }); }   // HLSL returns int; kept in E so that ?: arms agree
template <class A, HLSL_NUM(A)> inline auto isnan(const A& a) { return build<bool, Tr<A>::N>([&](int i) { return std::isnan((float)Tr<A>::get(a, i)); }); }
template <class A, HLSL_NUM(A)> inline auto isinf(const A& a) { return build<bool, Tr<A>::N>([&](int i) { return std::isinf((float)Tr<A>::get(a, i)); }); }
template <class A, HLSL_NUM(A)> inline auto isfinite(const A& a) { return build<bool, Tr<A>::N>([&](int i) { return std::isfinite((float)Tr<A>::get(a, i)); }); }
template <class A, HLSL_NUM(A)> inline bool any(const A& a) { bool r = false; for (int i = 0; i < Tr<A>::N; i++) r = r || (Tr<A>::get(a, i) != 0); return r; }
template <class A, HLSL_NUM(A)> inline bool all(const A& a) { bool r = true; for (int i = 0; i < Tr<A>::N; i++) r = r && (Tr<A>::get(a, i) != 0); return r; }
template <class A, class B, HLSL_NUM2(A, B)> inline auto dot(const A& a, const B& b) {
    typedef Prom<A, B> E; constexpr int N = maxN<A, B>(); E s = (E)Tr<A>::get(a, 0) * (E)Tr<B>::get(b, 0);
    for (int i = 1; i < N; i++) s = s + (E)Tr<A>::get(a, i) * (E)Tr<B>::get(b, i);
    return s; }
template <class A, HLSL_NUM(A)> inline float length(const A& a) { return std::sqrt((float)dot(a, a)); }
template <class A, class B, HLSL_NUM2(A, B)> inline float distance(const A& a, const B& b) { return length(a - b); }
template <class A, HLSL_NUM(A)> inline auto normalize(const A& a) { return a * (1.0f / length(a)); }
template <class A, class B, HLSL_NUM2(A, B)> inline float3 cross(const A& a_, const B& b_) { float3 a(a_), b(b_); return float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
template <class A, class B, HLSL_NUM2(A, B)> inline auto reflect(const A& i, const B& n) { return i - 2.0f * dot(i, n) * n; }
template <class S, class C> inline void sincos(float x, S& s, C& c) { s = std::sin(x); c = std::cos(x); }

inline uint bitsOf(float f) { uint u; memcpy(&u, &f, 4); return u; }
inline float floatOf(uint u) { float f; memcpy(&f, &u, 4); return f; }
template <class A, HLSL_NUM(A)> inline auto asuint(const A& a) {
    return build<uint, Tr<A>::N>([&](int i) { auto x = Tr<A>::get(a, i); if constexpr (std::is_floating_point_v<decltype(x)>) return bitsOf(x); else return (uint)x; }); }
template <class A, HLSL_NUM(A)> inline auto asint(const A& a) {
    return build<int, Tr<A>::N>([&](int i) { auto x = Tr<A>::get(a, i); if constexpr (std::is_floating_point_v<decltype(x)>) return (int)bitsOf(x); else return (int)x; }); }
template <class A, HLSL_NUM(A)> inline auto asfloat(const A& a) {
    return build<float, Tr<A>::N>([&](int i) { auto x = Tr<A>::get(a, i); if constexpr (std::is_floating_point_v<decltype(x)>) return (float)x; else return floatOf((uint)x); }); }

// fp16 <-> fp32 (round to nearest even, denormals kept)
inline uint halfBits(float f) {
    uint x = bitsOf(f), sgn = (x >> 16) & 0x8000u, a = x & 0x7FFFFFFFu;
    if (a > 0x7F800000u) return sgn | 0x7E00u;             // NaN
    if (a >= 0x477FF000u) return sgn | 0x7C00u;            // >= 65520 rounds to inf
    if (a < 0x33000000u) return sgn;                       // < 2^-25 rounds to 0 ( exactly 2^-25 ties to even = 0 )
    int e = (int)(a >> 23) - 127;
    uint m = (a & 0x7FFFFFu) | 0x800000u;                  // 24-bit significand
    int shift = e >= -14 ? 13 : 13 + (-14 - e);            // bits dropped
    uint h = m >> shift, rem = m & ((1u << shift) - 1u), half = 1u << (shift - 1);
    if (rem > half || (rem == half && (h & 1u))) h++;
    // normal: h has the implicit bit at 0x400 -> exponent field adds ( e + 14 ) << 10; subnormal: h is the mantissa itself ( may carry into the exponent )
    uint r = e >= -14 ? ((uint)(e + 14) << 10) + h : h;
    return sgn | r;
}
inline float halfToFloat(uint h) {
    uint s = (h & 0x8000u) << 16, e = (h >> 10) & 31u, m = h & 1023u;
    if (e == 0) { if (!m) return floatOf(s); float f = (float)m * 5.9604644775390625e-8f; return (h & 0x8000u) ? -f : f; }
    if (e == 31) return floatOf(s | 0x7F800000u | (m << 13));
    return floatOf(s | ((e + 112u) << 23) | (m << 13));
}
template <class A, HLSL_NUM(A)> inline auto f32tof16(const A& a) { return build<uint, Tr<A>::N>([&](int i) { return halfBits((float)Tr<A>::get(a, i)); }); }
template <class A, HLSL_NUM(A)> inline auto f16tof32(const A& a) { return build<float, Tr<A>::N>([&](int i) { return halfToFloat((uint)Tr<A>::get(a, i) & 0xFFFFu); }); }
template <class A, HLSL_NUM(A)> inline auto countbits(const A& a) { return build<uint, Tr<A>::N>([&](int i) { return (uint)__builtin_popcount((uint)Tr<A>::get(a, i)); }); }
template <class A, HLSL_NUM(A)> inline auto reversebits(const A& a) {
    return build<uint, Tr<A>::N>([&](int i) { uint x = (uint)Tr<A>::get(a, i), r = 0; for (int k = 0; k < 32; k++) r |= ((x >> k) & 1u) << (31 - k); return r; }); }
template <class A, HLSL_NUM(A)> inline auto firstbithigh(const A& a) { return build<uint, Tr<A>::N>([&](int i) { uint x = (uint)Tr<A>::get(a, i); return x ? 31u - (uint)__builtin_clz(x) : 0xFFFFFFFFu; }); }
template <class A, HLSL_NUM(A)> inline auto firstbitlow(const A& a) { return build<uint, Tr<A>::N>([&](int i) { uint x = (uint)Tr<A>::get(a, i); return x ? (uint)__builtin_ctz(x) : 0xFFFFFFFFu; }); }

// scalar "swizzles" ( s.xxx ): hlsl2cpp.py rewrites IDENT.xx / .xxx / .xxxx into these
template <class A, HLSL_NUM(A)> inline auto hlsl_splat2(const A& a) { typedef typename Tr<A>::E E; return vec<E, 2>(Tr<A>::get(a, 0), Tr<A>::get(a, 0)); }
template <class A, HLSL_NUM(A)> inline auto hlsl_splat3(const A& a) { typedef typename Tr<A>::E E; return vec<E, 3>(Tr<A>::get(a, 0), Tr<A>::get(a, 0), Tr<A>::get(a, 0)); }
template <class A, HLSL_NUM(A)> inline auto hlsl_splat4(const A& a) { typedef typename Tr<A>::E E; return vec<E, 4>(Tr<A>::get(a, 0), Tr<A>::get(a, 0), Tr<A>::get(a, 0), Tr<A>::get(a, 0)); }
// vector ?: ( HLSL before 2021 ): hlsl2cpp.py can rewrite a listed expression into this
template <class C, class A, class B> inline auto hlsl_select(const C& c, const A& a, const B& b) {
    typedef Prom<A, B> E; constexpr int N = maxN<RT<float, maxN<A, B>()>, C>();
    return build<E, N>([&](int i) { return Tr<C>::get(c, i) ? (E)Tr<A>::get(a, i) : (E)Tr<B>::get(b, i); }); }

// ------------------------------------------------------------------------------------------------ matrices (row vectors; m[ r ][ c ])
template <int R, int C> struct mat {
    vec<float, C> r[R];
    mat() {}
    template <class A0, class... AR, std::enable_if_t<(Tr<A0>::num && (... && Tr<AR>::num)) && (Tr<A0>::N + (0 + ... + Tr<AR>::N) == R * C) && (sizeof...(AR) > 0), int> = 0>
    mat(const A0& a0, const AR&... ar) { float t[R * C]; int k = 0; flattenInto<float, R * C>(t, k, a0); (flattenInto<float, R * C>(t, k, ar), ...);
        for (int i = 0; i < R; i++) for (int j = 0; j < C; j++) r[i].d[j] = t[i * C + j]; }
    template <int R2, int C2, std::enable_if_t<(R2 >= R && C2 >= C && (R2 > R || C2 > C)), int> = 0> explicit mat(const mat<R2, C2>& m) {
        for (int i = 0; i < R; i++) for (int j = 0; j < C; j++) r[i].d[j] = m.r[i].d[j]; }
    vec<float, C>& operator[](int i) { return r[i]; }
    const vec<float, C>& operator[](int i) const { return r[i]; }
};
typedef mat<2, 2> float2x2; typedef mat<3, 3> float3x3; typedef mat<4, 4> float4x4; typedef mat<2, 3> float2x3; typedef mat<3, 4> float3x4; typedef mat<3, 2> float3x2; typedef mat<4, 3> float4x3;
template <int R, int C> inline mat<C, R> transpose(const mat<R, C>& m) { mat<C, R> t; for (int i = 0; i < R; i++) for (int j = 0; j < C; j++) t.r[j].d[i] = m.r[i].d[j]; return t; }
// mul( M, v ): v is a column; mul( v, M ): v is a row
template <int R, int C, class V, class = std::enable_if_t<Tr<V>::hl && Tr<V>::N == C>> inline vec<float, R> mul(const mat<R, C>& m, const V& v_) {
    vec<float, C> v(v_); vec<float, R> o; for (int i = 0; i < R; i++) o.d[i] = dot(m.r[i], v); return o; }
template <int R, int C, class V, class = std::enable_if_t<Tr<V>::hl && Tr<V>::N == R>> inline vec<float, C> mul(const V& v_, const mat<R, C>& m) {
    vec<float, R> v(v_); vec<float, C> o; for (int j = 0; j < C; j++) { float s = v.d[0] * m.r[0].d[j]; for (int i = 1; i < R; i++) s += v.d[i] * m.r[i].d[j]; o.d[j] = s; } return o; }
template <int R, int K, int C> inline mat<R, C> mul(const mat<R, K>& a, const mat<K, C>& b) {
    mat<R, C> o; for (int i = 0; i < R; i++) for (int j = 0; j < C; j++) { float s = a.r[i].d[0] * b.r[0].d[j]; for (int k = 1; k < K; k++) s += a.r[i].d[k] * b.r[k].d[j]; o.r[i].d[j] = s; } return o; }
template <class A, class B, class = std::enable_if_t<Tr<A>::num && Tr<B>::num>> inline auto mul(const A& a, const B& b) { if constexpr (Tr<A>::N > 1 && Tr<B>::N > 1) return dot(a, b); else return a * b; }

// ------------------------------------------------------------------------------------------------ textures
enum Format : uint32_t {  // numeric values = nrd::Format ( NRDDescs.h )
    R8_UNORM = 0, R8_UINT = 2, RG8_UNORM = 4, RGBA8_UNORM = 8, R16_UNORM = 13, R16_UINT = 15, R16_SFLOAT = 17, RG16_SFLOAT = 22, RGBA16_SNORM = 24, RGBA16_SFLOAT = 27, R32_UINT = 28, R32_SFLOAT = 30, RGBA32_SFLOAT = 39, R10_G10_B10_A2_UNORM = 40,
};
struct HostTexture { void* data; uint32_t width, height, pitchBytes, format; };   // same layout as the oracle's OracleTexture

struct TexData {
    uint8_t* data = nullptr; int w = 0, h = 0, pitch = 0; uint32_t fmt = 0;
    bool inside(int x, int y) const { return x >= 0 && y >= 0 && x < w && y < h; }
    int bpp() const { switch (fmt) { case R8_UNORM: case R8_UINT: return 1; case RG8_UNORM: case R16_UINT: case R16_SFLOAT: case R16_UNORM: return 2; case RGBA16_SFLOAT: case RGBA16_SNORM: return 8; case RGBA32_SFLOAT: return 16; default: return 4; } }
    uint8_t* at(int x, int y) const { return data + (size_t)y * pitch + (size_t)x * bpp(); }
    float4 fetch(int x, int y) const {
        const uint8_t* p = at(x, y);
        switch (fmt) {
            case R8_UNORM: return float4(p[0] / 255.0f, 0, 0, 1);
            case RG8_UNORM: return float4(p[0] / 255.0f, p[1] / 255.0f, 0, 1);
            case RGBA8_UNORM: return float4(p[0] / 255.0f, p[1] / 255.0f, p[2] / 255.0f, p[3] / 255.0f);
            case R16_SFLOAT: { uint16_t v; memcpy(&v, p, 2); return float4(halfToFloat(v), 0, 0, 1); }
            case R16_UNORM: { uint16_t v; memcpy(&v, p, 2); return float4(v / 65535.0f, 0, 0, 1); }
            case RGBA16_SNORM: { int16_t v[4]; memcpy(v, p, 8); return float4(std::fmax(v[0] / 32767.0f, -1.0f), std::fmax(v[1] / 32767.0f, -1.0f), std::fmax(v[2] / 32767.0f, -1.0f), std::fmax(v[3] / 32767.0f, -1.0f)); }
            case RG16_SFLOAT: { uint16_t v[2]; memcpy(v, p, 4); return float4(halfToFloat(v[0]), halfToFloat(v[1]), 0, 1); }
            case RGBA16_SFLOAT: { uint16_t v[4]; memcpy(v, p, 8); return float4(halfToFloat(v[0]), halfToFloat(v[1]), halfToFloat(v[2]), halfToFloat(v[3])); }
            case R32_SFLOAT: { float v; memcpy(&v, p, 4); return float4(v, 0, 0, 1); }
            case RGBA32_SFLOAT: { float v[4]; memcpy(v, p, 16); return float4(v[0], v[1], v[2], v[3]); }
            case R10_G10_B10_A2_UNORM: { uint32_t v; memcpy(&v, p, 4); return float4((v & 1023u) / 1023.0f, ((v >> 10) & 1023u) / 1023.0f, ((v >> 20) & 1023u) / 1023.0f, (v >> 30) / 3.0f); }
            case R8_UINT: return float4(floatOf((uint32_t)p[0]), 0, 0, 0);
            case R16_UINT: { uint16_t v; memcpy(&v, p, 2); return float4(floatOf(v), 0, 0, 0); }      // raw bits travel in .x for <uint> views
            case R32_UINT: { uint32_t v; memcpy(&v, p, 4); return float4(floatOf(v), 0, 0, 0); }
            default: return float4(0.0f);
        }
    }
    static uint32_t unorm(float v, float m) { float s = std::fmin(std::fmax(v, 0.0f), 1.0f); return (uint32_t)(s * m + 0.5f); }
    void store(int x, int y, const float4& v) const {
        if (!inside(x, y)) return;
        uint8_t* p = at(x, y);
        switch (fmt) {
            case R8_UNORM: p[0] = (uint8_t)unorm(v.x, 255.0f); break;
            case RG8_UNORM: p[0] = (uint8_t)unorm(v.x, 255.0f); p[1] = (uint8_t)unorm(v.y, 255.0f); break;
            case RGBA8_UNORM: for (int i = 0; i < 4; i++) p[i] = (uint8_t)unorm(v.d[i], 255.0f); break;
            case R16_SFLOAT: { uint16_t q = (uint16_t)halfBits(v.x); memcpy(p, &q, 2); break; }
            case R16_UNORM: { uint16_t q = (uint16_t)unorm(v.x, 65535.0f); memcpy(p, &q, 2); break; }
            case RGBA16_SNORM: { int16_t q[4]; for (int i = 0; i < 4; i++) { float s = std::fmin(std::fmax(v.d[i], -1.0f), 1.0f) * 32767.0f; q[i] = (int16_t)(int)(s + (s >= 0.0f ? 0.5f : -0.5f)); } memcpy(p, q, 8); break; }
            case RG16_SFLOAT: { uint16_t q[2] = {(uint16_t)halfBits(v.x), (uint16_t)halfBits(v.y)}; memcpy(p, q, 4); break; }
            case RGBA16_SFLOAT: { uint16_t q[4] = {(uint16_t)halfBits(v.x), (uint16_t)halfBits(v.y), (uint16_t)halfBits(v.z), (uint16_t)halfBits(v.w)}; memcpy(p, q, 8); break; }
            case R32_SFLOAT: memcpy(p, &v.x, 4); break;
            case RGBA32_SFLOAT: { float q[4] = {v.x, v.y, v.z, v.w}; memcpy(p, q, 16); break; }
            case R10_G10_B10_A2_UNORM: { uint32_t q = unorm(v.x, 1023.0f) | (unorm(v.y, 1023.0f) << 10) | (unorm(v.z, 1023.0f) << 20) | (unorm(v.w, 3.0f) << 30); memcpy(p, &q, 4); break; }
            case R8_UINT: { uint32_t q = bitsOf(v.x); p[0] = (uint8_t)(q > 255u ? 255u : q); break; }   // D3D clamps integer stores to the format's range
            case R16_UINT: { uint16_t q = (uint16_t)bitsOf(v.x); memcpy(p, &q, 2); break; }
            case R32_UINT: { uint32_t q = bitsOf(v.x); memcpy(p, &q, 4); break; }
            default: break;
        }
    }
    float4 load(int x, int y) const { return inside(x, y) ? fetch(x, y) : float4(0.0f); }
    float4 fetchClamped(int x, int y) const { return fetch(x < 0 ? 0 : (x >= w ? w - 1 : x), y < 0 ? 0 : (y >= h ? h - 1 : y)); }
};

struct SamplerState { bool linear = false; SamplerState() {} explicit SamplerState(const char* name) : linear(strstr(name, "inear") != nullptr) {} };

// float4 <-> the declared element type of a view
template <class T> struct Elem;
template <> struct Elem<float> { typedef Scalar1<float> R; static float from(const float4& v) { return v.x; } static float4 to(float v) { return float4(v, 0, 0, 0); } typedef float4 G; static G gather(const float* c) { return G(c[0], c[1], c[2], c[3]); } };
template <> struct Elem<float2> { typedef float2 R; static float2 from(const float4& v) { return v.xy; } static float4 to(const float2& v) { return float4(v, 0, 0); } typedef float4 G; static G gather(const float* c) { return G(c[0], c[1], c[2], c[3]); } };
template <> struct Elem<float3> { typedef float3 R; static float3 from(const float4& v) { return v.xyz; } static float4 to(const float3& v) { return float4(v, 0); } typedef float4 G; static G gather(const float* c) { return G(c[0], c[1], c[2], c[3]); } };
template <> struct Elem<float4> { typedef float4 R; static float4 from(const float4& v) { return v; } static float4 to(const float4& v) { return v; } typedef float4 G; static G gather(const float* c) { return G(c[0], c[1], c[2], c[3]); } };
template <> struct Elem<uint> { typedef Scalar1<uint> R; static uint from(const float4& v) { return bitsOf(v.x); } static float4 to(uint v) { return float4(floatOf(v), 0, 0, 0); } typedef uint4 G; static G gather(const float* c) { return G(bitsOf(c[0]), bitsOf(c[1]), bitsOf(c[2]), bitsOf(c[3])); } };
template <> struct Elem<uint2> { typedef uint2 R; static uint2 from(const float4& v) { return uint2(bitsOf(v.x), bitsOf(v.y)); } static float4 to(const uint2& v) { return float4(floatOf(v.x), floatOf(v.y), 0, 0); } typedef uint4 G; static G gather(const float* c) { return G(bitsOf(c[0]), bitsOf(c[1]), bitsOf(c[2]), bitsOf(c[3])); } };

template <> struct Elem<uint4> { typedef uint4 R; static uint4 from(const float4& v) { return uint4(bitsOf(v.x), bitsOf(v.y), bitsOf(v.z), bitsOf(v.w)); } static float4 to(const uint4& v) { return float4(floatOf(v.x), floatOf(v.y), floatOf(v.z), floatOf(v.w)); } typedef uint4 G; static G gather(const float* c) { return G(bitsOf(c[0]), bitsOf(c[1]), bitsOf(c[2]), bitsOf(c[3])); } };

template <class T> struct Texture2D {
    TexData t;
    template <class P> typename Elem<T>::R operator[](const P& p_) const { int2 p(p_); return typename Elem<T>::R(Elem<T>::from(t.load(p.x, p.y))); }
    template <class P> typename Elem<T>::R Load(const P& p_) const { vec<int, Tr<P>::N> p(p_); return typename Elem<T>::R(Elem<T>::from(t.load(p.d[0], p.d[1]))); }
    template <class P, class O> typename Elem<T>::R Load(const P& p_, const O& o_) const { vec<int, Tr<P>::N> p(p_); int2 o(o_); return typename Elem<T>::R(Elem<T>::from(t.load(p.d[0] + o.x, p.d[1] + o.y))); }
    float4 sample(const SamplerState& s, const float2& uv, const int2& off) const {
        if (!s.linear) return t.fetchClamped((int)std::floor(uv.x * (float)t.w) + off.x, (int)std::floor(uv.y * (float)t.h) + off.y);
        float tx = uv.x * (float)t.w - 0.5f, ty = uv.y * (float)t.h - 0.5f, fx = std::floor(tx), fy = std::floor(ty), wx = tx - fx, wy = ty - fy;
        int x0 = (int)fx + off.x, y0 = (int)fy + off.y;
        float4 a = t.fetchClamped(x0, y0), b = t.fetchClamped(x0 + 1, y0), c = t.fetchClamped(x0, y0 + 1), d = t.fetchClamped(x0 + 1, y0 + 1);
        float4 top = a + (b - a) * wx, bot = c + (d - c) * wx;
        return float4(top + (bot - top) * wy);
    }
    template <class U, class L> typename Elem<T>::R SampleLevel(const SamplerState& s, const U& uv, const L&) const { return typename Elem<T>::R(Elem<T>::from(sample(s, float2(uv), int2(0, 0)))); }
    template <class U, class L, class O> typename Elem<T>::R SampleLevel(const SamplerState& s, const U& uv, const L&, const O& off) const { return typename Elem<T>::R(Elem<T>::from(sample(s, float2(uv), int2(off)))); }
    template <class U> typename Elem<T>::G gatherChannel(const U& uv_, int ch, const int2& off) const {
        float2 uv(uv_);
        int x0 = (int)std::floor(uv.x * (float)t.w - 0.5f) + off.x, y0 = (int)std::floor(uv.y * (float)t.h - 0.5f) + off.y;
        float c[4] = {t.fetchClamped(x0, y0 + 1).d[ch], t.fetchClamped(x0 + 1, y0 + 1).d[ch], t.fetchClamped(x0 + 1, y0).d[ch], t.fetchClamped(x0, y0).d[ch]};
        return Elem<T>::gather(c);
    }
    template <class U> auto GatherRed(const SamplerState&, const U& uv) const { return gatherChannel(uv, 0, int2(0, 0)); }
    template <class U> auto GatherGreen(const SamplerState&, const U& uv) const { return gatherChannel(uv, 1, int2(0, 0)); }
    template <class U> auto GatherBlue(const SamplerState&, const U& uv) const { return gatherChannel(uv, 2, int2(0, 0)); }
    template <class U> auto GatherAlpha(const SamplerState&, const U& uv) const { return gatherChannel(uv, 3, int2(0, 0)); }
    template <class U, class O> auto GatherRed(const SamplerState&, const U& uv, const O& o) const { return gatherChannel(uv, 0, int2(o)); }
    template <class U, class O> auto GatherGreen(const SamplerState&, const U& uv, const O& o) const { return gatherChannel(uv, 1, int2(o)); }
    template <class A, class B> void GetDimensions(A& w, B& h) const { w = (A)t.w; h = (B)t.h; }
};
template <class T> struct RWRef {
    const TexData* t; int x, y;
    operator T() const { return Elem<T>::from(t->load(x, y)); }
    template <class V> const RWRef& operator=(const V& v) const { t->store(x, y, Elem<T>::to(T(v))); return *this; }
    const RWRef& operator=(const RWRef& o) const { return *this = (T)o; }
};
template <class T> struct RWTexture2D {
    TexData t;
    template <class P> RWRef<T> operator[](const P& p_) const { int2 p(p_); return RWRef<T>{&t, p.x, p.y}; }
    template <class A, class B> void GetDimensions(A& w, B& h) const { w = (A)t.w; h = (B)t.h; }
};

// ------------------------------------------------------------------------------------------------ per-shader registry
struct ConstReg { void* p; int kind; };   // kind: number of 32-bit words ( 1..4 ), 16 = float4x4
struct SmemReg { void* (*address)(); size_t bytes; };   // a groupshared array: thread_local, so its address is asked for on the executing thread
struct ShaderModule {
    std::vector<ConstReg> constants;
    std::vector<TexData*> srv, uav;
    std::vector<SmemReg> smem;
};
// Group-shared memory starts every group zero-filled. ( On a GPU it starts with leftovers: the RELAX / REBLUR shaders skip their preload on sky
// tiles and still read it — e.g. RELAX_AtrousSmem.cs.hlsl:145-150 writes gOut_MaterialID from it — so those texels are garbage by design;
// zero makes them reproducible. )
struct SmemRegistrar { SmemRegistrar(ShaderModule* m, void* (*address)(), size_t bytes) { m->smem.push_back({address, bytes}); } };
struct ConstRegistrar {
    template <class T> ConstRegistrar(ShaderModule* m, T* p) {
        int kind = 0;
        if constexpr (std::is_same_v<T, float4x4>) kind = 16; else if constexpr (Tr<T>::num) kind = Tr<T>::N;
        static_assert(std::is_same_v<T, float4x4> || Tr<T>::num, "unsupported constant type");
        m->constants.push_back({p, kind});
    }
};
struct TexRegistrar {
    TexRegistrar(ShaderModule* m, TexData* t, char reg, int index) {
        auto& v = reg == 't' ? m->srv : m->uav;
        if ((int)v.size() <= index) v.resize(index + 1, nullptr);
        v[index] = t;
    }
};
// HLSL constant-buffer packing: 4-byte scalars, a vector never straddles a 16-byte register, matrices are 4 registers stored by column
// ( #pragma pack_matrix( column_major ), NRD.hlsli:326 )
inline bool loadConstants(const ShaderModule& m, const void* data, uint32_t size) {
    const uint8_t* b = (const uint8_t*)data; uint32_t off = 0;
    for (const ConstReg& c : m.constants) {
        uint32_t bytes = (c.kind == 16 ? 64u : (uint32_t)c.kind * 4u);
        if (c.kind == 16) off = (off + 15u) & ~15u;
        else if ((off & 15u) + bytes > 16u) off = (off + 15u) & ~15u;
        if (off + bytes > size) return false;
        if (c.kind == 16) { float t[16]; memcpy(t, b + off, 64); float4x4* M = (float4x4*)c.p; for (int r = 0; r < 4; r++) for (int col = 0; col < 4; col++) M->r[r].d[col] = t[col * 4 + r]; }
        else memcpy(c.p, b + off, bytes);
        off += bytes;
    }
    return ((off + 15u) & ~15u) == ((size + 15u) & ~15u);
}

// ------------------------------------------------------------------------------------------------ thread groups as fibers
struct Fiber {
    ucontext_t ctx; char* stack = nullptr; int state = 0;   // 0 runnable, 1 at barrier, 2 done
    uint3 groupThreadID; uint groupIndex = 0;
    uint8_t xchg[64]; uint32_t xchgSeq = 0;
    uint rngHashState = 0; uint2 rngTeaState;
    uint32_t probeCalls = 0;   // probeMirror calls of this thread so far: taps 0-7 are the first lobe's, 8-15 the second's
};
struct GroupRun {
    std::vector<Fiber> fibers; ucontext_t sched; int current = -1; bool useFibers = false;
    void (*entry)(uint3, uint3, uint3, uint) = nullptr; uint3 groupID; uint3 groupSize;
};
GroupRun& groupRun();
inline Fiber& currentFiber() { GroupRun& g = groupRun(); return g.fibers[g.current < 0 ? 0 : g.current]; }
inline void fiberYield(int state) { GroupRun& g = groupRun(); Fiber& f = g.fibers[g.current]; f.state = state; swapcontext(&f.ctx, &g.sched); }
inline void GroupMemoryBarrierWithGroupSync() { if (groupRun().useFibers) fiberYield(1); }
// SIGMA_ClassifyTiles relies on a 32-thread group ( one wave, lockstep ) + GroupMemoryBarrier(): on a CPU that is a sync point too
inline void GroupMemoryBarrier() { GroupMemoryBarrierWithGroupSync(); }
inline void AllMemoryBarrierWithGroupSync() { GroupMemoryBarrierWithGroupSync(); }
template <class A> inline A quadRead(const A& v, uint mask) {
    GroupRun& g = groupRun();
    if (!g.useFibers) { fprintf(stderr, "hlsl_cpu: quad read without fibers\n"); abort(); }
    static_assert(sizeof(A) <= 64, "quad exchange payload");
    Fiber& f = g.fibers[g.current];
    memcpy(f.xchg, &v, sizeof(A)); f.xchgSeq++;
    const uint32_t seq = f.xchgSeq;
    fiberYield(0);
    A r = v;
    Fiber& p = g.fibers[f.groupIndex ^ mask];
    if (p.xchgSeq == seq && p.state != 2) memcpy((void*)&r, p.xchg, sizeof(A));
    fiberYield(0);
    return r;
}
// SM 6.0 quads in a compute shader: 4 consecutive lanes of the flattened group
template <class A> inline A QuadReadAcrossX(const A& v) { return quadRead(v, 1u); }
template <class A> inline A QuadReadAcrossY(const A& v) { return quadRead(v, 2u); }
template <class A> inline A QuadReadAcrossDiagonal(const A& v) { return quadRead(v, 3u); }
template <class D, class V> inline void InterlockedAdd(D& d, const V& v) { d = (D)(d + (D)v); }
template <class D, class V> inline void InterlockedMax(D& d, const V& v) { if ((D)v > d) d = (D)v; }
template <class D, class V> inline void InterlockedMin(D& d, const V& v) { if ((D)v < d) d = (D)v; }
template <class D, class V> inline void InterlockedOr(D& d, const V& v) { d = (D)(d | (D)v); }

void runGroup(const ShaderModule& module, void (*entry)(uint3, uint3, uint3, uint), uint3 groupID, uint3 groupSize, bool useFibers);

// what one compiled shader permutation exports
struct ShaderEntry {
    const char* identifier; ShaderModule* module; void (*entry)(uint3, uint3, uint3, uint); uint32_t gx, gy, gz; bool useFibers;
};
void registerShader(const ShaderEntry& e);
bool probeMirror(bool mirrored);   // hlsl_runtime.cpp: branch counter behind nrd_refshader_probe
}  // namespace hlsl
