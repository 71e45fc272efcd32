// TEST INFRASTRUCTURE — CPU oracle (see hlsl_like.h). Software texture unit.
//
// Semantics restated from what the reference's shaders rely on:
//  * Texture2D[pos] / Load: out-of-bounds reads return 0; RWTexture2D[pos] = v: out-of-bounds writes are dropped
//  * samplers are clamp-to-edge (NRDIntegration.hpp:395-406); bilinear weights are exact fp32 fractions
//    (D3D/VK hardware quantises them to 8 bits — documented tolerance source)
//  * GatherRed returns the 2x2 bilinear footprint, re-ordered here to (00, 10, 01, 11) == HLSL ".wzxy"
//  * stores quantise like the API formats: UNORM = floor(saturate(x) * max + 0.5), FP16 = round-to-nearest-even
// Pixel layout is pitch-linear, texel encodings exactly the nrd::Format ones (NRDDescs.h:264-322).
#pragma once
#include "hlsl_like.h"

namespace orc {

enum Fmt : uint32_t {  // numeric values = nrd::Format
    FMT_R8_UNORM = 0,
    FMT_R8_UINT = 2,
    FMT_RG8_UNORM = 4,
    FMT_RGBA8_UNORM = 8,
    FMT_R16_UINT = 15,
    FMT_R16_SFLOAT = 17,
    FMT_RG16_SFLOAT = 22,
    FMT_RGBA16_SFLOAT = 27,
    FMT_R32_UINT = 28,
    FMT_R32_SFLOAT = 30,
    FMT_RGBA32_SFLOAT = 39,
    FMT_R10_G10_B10_A2_UNORM = 40,
};

// C-ABI view handed in by the test driver (one per binding, in DispatchDesc::resources order)
struct OracleTexture {
    void* data;
    uint32_t width, height;
    uint32_t pitchBytes;
    uint32_t format;
};

inline uint32_t unormEncode(float v, float maxv) { return (uint32_t)(saturate(v) * maxv + 0.5f); }

struct Tex {
    uint8_t* data = nullptr;
    int w = 0, h = 0, pitch = 0;
    uint32_t fmt = 0;
    Tex() {}
    explicit Tex(const OracleTexture& t) : data((uint8_t*)t.data), w((int)t.width), h((int)t.height), pitch((int)t.pitchBytes), fmt(t.format) {}

    bool inside(int x, int y) const { return x >= 0 && y >= 0 && x < w && y < h; }
    const uint8_t* at(int x, int y, int bpp) const { return data + (size_t)y * pitch + (size_t)x * bpp; }
    uint8_t* at(int x, int y, int bpp) { return data + (size_t)y * pitch + (size_t)x * bpp; }

    // ---- typed fetch of an in-bounds texel ----
    float4 fetch(int x, int y) const {
        switch (fmt) {
            case FMT_R8_UNORM: return float4(*at(x, y, 1) / 255.0f, 0, 0, 1);
            case FMT_RG8_UNORM: { const uint8_t* p = at(x, y, 2); return float4(p[0] / 255.0f, p[1] / 255.0f, 0, 1); }
            case FMT_RGBA8_UNORM: { const uint8_t* p = at(x, y, 4); return float4(p[0] / 255.0f, p[1] / 255.0f, p[2] / 255.0f, p[3] / 255.0f); }
            case FMT_R16_SFLOAT: { uint16_t v; memcpy(&v, at(x, y, 2), 2); return float4(f16tof32(v), 0, 0, 1); }
            case FMT_RG16_SFLOAT: { uint16_t v[2]; memcpy(v, at(x, y, 4), 4); return float4(f16tof32(v[0]), f16tof32(v[1]), 0, 1); }
            case FMT_RGBA16_SFLOAT: { uint16_t v[4]; memcpy(v, at(x, y, 8), 8); return float4(f16tof32(v[0]), f16tof32(v[1]), f16tof32(v[2]), f16tof32(v[3])); }
            case FMT_R32_SFLOAT: { float v; memcpy(&v, at(x, y, 4), 4); return float4(v, 0, 0, 1); }
            case FMT_RGBA32_SFLOAT: { float v[4]; memcpy(v, at(x, y, 16), 16); return float4(v[0], v[1], v[2], v[3]); }
            case FMT_R10_G10_B10_A2_UNORM: {
                uint32_t v; memcpy(&v, at(x, y, 4), 4);
                return float4((v & 1023u) / 1023.0f, ((v >> 10) & 1023u) / 1023.0f, ((v >> 20) & 1023u) / 1023.0f, (v >> 30) / 3.0f);
            }
            default: return float4(0.0f);
        }
    }
    uint32_t fetchUint(int x, int y) const {
        if (fmt == FMT_R8_UINT) return *at(x, y, 1);
        if (fmt == FMT_R16_UINT) { uint16_t v; memcpy(&v, at(x, y, 2), 2); return v; }
        if (fmt == FMT_R32_UINT) { uint32_t v; memcpy(&v, at(x, y, 4), 4); return v; }
        return 0;
    }

    // ---- Texture2D[pos] ----
    float4 load(int x, int y) const { return inside(x, y) ? fetch(x, y) : float4(0.0f); }
    float4 load(int2 p) const { return load(p.x, p.y); }
    uint32_t loadUint(int x, int y) const { return inside(x, y) ? fetchUint(x, y) : 0u; }

    // ---- clamp-addressed fetch (what samplers see) ----
    float4 fetchClamped(int x, int y) const { return fetch(x < 0 ? 0 : (x >= w ? w - 1 : x), y < 0 ? 0 : (y >= h ? h - 1 : y)); }
    uint32_t fetchUintClamped(int x, int y) const { return fetchUint(x < 0 ? 0 : (x >= w ? w - 1 : x), y < 0 ? 0 : (y >= h ? h - 1 : y)); }

    // SampleLevel( gLinearClamp, uv, 0 )
    float4 sampleLinear(float2 uv) const {
        float tx = uv.x * (float)w - 0.5f, ty = uv.y * (float)h - 0.5f;
        float fx = std::floor(tx), fy = std::floor(ty);
        float wx = tx - fx, wy = ty - fy;
        int x0 = (int)fx, y0 = (int)fy;
        float4 a = fetchClamped(x0, y0), b = fetchClamped(x0 + 1, y0), c = fetchClamped(x0, y0 + 1), d = fetchClamped(x0 + 1, y0 + 1);
        return lerp(lerp(a, b, wx), lerp(c, d, wx), wy);
    }
    // SampleLevel( gNearestClamp, uv, 0 )
    float4 sampleNearest(float2 uv) const { return fetchClamped((int)std::floor(uv.x * (float)w), (int)std::floor(uv.y * (float)h)); }

    // ---- RWTexture2D[pos] = v ----
    void store(int x, int y, float4 v) {
        if (!inside(x, y)) return;
        switch (fmt) {
            case FMT_R8_UNORM: *at(x, y, 1) = (uint8_t)unormEncode(v.x, 255.0f); break;
            case FMT_RG8_UNORM: { uint8_t* p = at(x, y, 2); p[0] = (uint8_t)unormEncode(v.x, 255.0f); p[1] = (uint8_t)unormEncode(v.y, 255.0f); break; }
            case FMT_RGBA8_UNORM: { uint8_t* p = at(x, y, 4); for (int i = 0; i < 4; i++) p[i] = (uint8_t)unormEncode(v[i], 255.0f); break; }
            case FMT_R16_SFLOAT: { uint16_t h16 = f32tof16(v.x); memcpy(at(x, y, 2), &h16, 2); break; }
            case FMT_RG16_SFLOAT: { uint16_t q[2] = {f32tof16(v.x), f32tof16(v.y)}; memcpy(at(x, y, 4), q, 4); break; }
            case FMT_RGBA16_SFLOAT: { uint16_t q[4] = {f32tof16(v.x), f32tof16(v.y), f32tof16(v.z), f32tof16(v.w)}; memcpy(at(x, y, 8), q, 8); break; }
            case FMT_R32_SFLOAT: memcpy(at(x, y, 4), &v.x, 4); break;
            case FMT_RGBA32_SFLOAT: { float q[4] = {v.x, v.y, v.z, v.w}; memcpy(at(x, y, 16), q, 16); break; }
            case FMT_R10_G10_B10_A2_UNORM: {
                uint32_t q = unormEncode(v.x, 1023.0f) | (unormEncode(v.y, 1023.0f) << 10) | (unormEncode(v.z, 1023.0f) << 20) | (unormEncode(v.w, 3.0f) << 30);
                memcpy(at(x, y, 4), &q, 4);
                break;
            }
            default: break;
        }
    }
    void store(int2 p, float4 v) { store(p.x, p.y, v); }
    void storeUint(int x, int y, uint32_t v) {
        if (!inside(x, y)) return;
        if (fmt == FMT_R8_UINT) *at(x, y, 1) = (uint8_t)(v > 255u ? 255u : v);  // D3D clamps integer stores to the format's range
        else if (fmt == FMT_R16_UINT) { uint16_t q = (uint16_t)v; memcpy(at(x, y, 2), &q, 2); }
        else if (fmt == FMT_R32_UINT) memcpy(at(x, y, 4), &v, 4);
    }
    // raw copy of one texel between same-format textures (used where a shader forwards a packed value untouched)
    int bytesPerTexel() const {
        switch (fmt) {
            case FMT_R8_UNORM: case FMT_R8_UINT: return 1;
            case FMT_RG8_UNORM: case FMT_R16_UINT: case FMT_R16_SFLOAT: return 2;
            case FMT_RGBA16_SFLOAT: return 8;
            case FMT_RGBA32_SFLOAT: return 16;
            default: return 4;
        }
    }
};

}  // namespace orc
