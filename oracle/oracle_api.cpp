// TEST INFRASTRUCTURE — CPU oracle (see hlsl_like.h). C entry point used by tests/, smoke() and bench.py's CPU legs.
// One call replays one nrd::DispatchDesc on host memory: the shader is selected by the same `shaderIdentifier`
// string the product's CUDA executor keys on, textures arrive in DispatchDesc::resources order.
#include <omp.h>

#include <cstdio>
#include <cstring>
#include <string>

#include "reblur_passes.h"

namespace orc {
// sigma_passes.cpp
int sigmaDispatch(const std::string& id, const void* cb, uint32_t cbSize, Tex* t, uint32_t n, int gridW, int gridH);
// relax_passes.cpp
int relaxDispatch(const std::string& id, const void* cb, uint32_t cbSize, Tex* t, uint32_t n, int gridW, int gridH);
}  // namespace orc

using namespace orc;

// REFERENCE_TemporalAccumulation.cs.hlsl:18-28, REFERENCE_Copy.cs.hlsl:18-27 (16x16 groups; constants: REFERENCE_*.resources.hlsli:10-18)
static int referenceDispatch(const std::string& id, const void* constants, uint32_t cbSize, Tex* t, uint32_t n, int gridW, int gridH) {
    if (n != 2) return 2;
    const float* c = (const float*)constants;
    if (id == "REFERENCE_TemporalAccumulation.cs.hlsl") {
        if (cbSize != 16) return 2;
        const float gAccumSpeed = c[0];
        for (int y = 0; y < gridH * 16; y++)
            for (int x = 0; x < gridW * 16; x++) {
                float4 input = t[0].load(x, y), history = t[1].load(x, y);
                t[1].store(x, y, lerp(history, input, gAccumSpeed));
            }
        return 0;
    }
    if (id == "REFERENCE_Copy.cs.hlsl") {
        if (cbSize != 24) return 2;
        const float2 gRectSizeInv = float2(c[0], c[1]);
        const float gSplitScreen = c[2];
        for (int y = 0; y < gridH * 16; y++)
            for (int x = 0; x < gridW * 16; x++) {
                float2 pixelUv = float2(x + 0.5f, y + 0.5f) * gRectSizeInv;
                if (pixelUv.x > gSplitScreen) t[1].store(x, y, t[0].load(x, y));
            }
        return 0;
    }
    return 1;
}

// REBLUR_SplitScreen.cs.hlsl:21-56 (NRD_MODE = RADIANCE): the noisy input left of the split line
static void reblurSplitScreen(const ReblurCB& cb, const Tex& gIn_ViewZ, const Tex& gIn_Diff, const Tex& gIn_Spec, Tex& gOut_Diff, Tex& gOut_Spec, int gridW, int gridH, int signal) {
    ReblurCtx c(cb, signal);
    for (int py = 0; py < gridH * 16; py++)
        for (int px = 0; px < gridW * 8; px++) {
            float2 pixelUv = float2(px + 0.5f, py + 0.5f) * cb.gRectSizeInv;
            if (pixelUv.x > cb.gSplitScreen || px > cb.gRectSizeMinusOne.x || py > cb.gRectSizeMinusOne.y) continue;
            float viewZ = c.UnpackViewZ(gIn_ViewZ.load(px, py).x);
            float inRange = float(c.IsInDenoisingRange(viewZ));
            if (c.hasDiff()) gOut_Diff.store(px, py, gIn_Diff.load(px >> (cb.gDiffCheckerboard != 2 ? 1 : 0), py) * inRange);
            if (c.hasSpec()) gOut_Spec.store(px, py, gIn_Spec.load(px >> (cb.gSpecCheckerboard != 2 ? 1 : 0), py) * inRange);
        }
}

static bool startsWith(const std::string& s, const char* p) { return s.compare(0, strlen(p), p) == 0; }

extern "C" {

// flags: bit0 = replay SM6.0 quad intrinsics (NRD_SUPPORTS_QUAD_INTRINSICS = 1, the reference default)
//        bit1 = robust "tap left the screen" test instead of the bit-fragile any( uv != MirrorUv( uv ) ) (strict parity runs)
// returns 0 on success, 1 unknown shader, 2 bad arguments
__attribute__((visibility("default"))) int nrd_oracle_dispatch(const char* shaderIdentifier, const void* constants, uint32_t constantsSize,
                                                                const OracleTexture* textures, uint32_t texturesNum, uint32_t gridW, uint32_t gridH,
                                                                uint32_t flags) {
    if (!shaderIdentifier || (!textures && texturesNum)) return 2;
    std::string id = shaderIdentifier;
    Tex t[40];
    if (texturesNum > 40) return 2;
    for (uint32_t i = 0; i < texturesNum; i++) t[i] = Tex(textures[i]);
    const bool quads = flags & 1u, robust = flags & 2u;
    const int gw = (int)gridW, gh = (int)gridH;

    if (startsWith(id, "Clear.cs.hlsl")) {
        if (texturesNum != 1) return 2;
        clearTexture(t[0]);
        return 0;
    }
    if (startsWith(id, "REBLUR_")) {
        if (constantsSize != sizeof(ReblurCB)) return 2;
        const ReblurCB& cb = *(const ReblurCB*)constants;
        if (startsWith(id, "REBLUR_ClassifyTiles.cs.hlsl")) {
            if (texturesNum != 2) return 2;
            reblurClassifyTiles(cb, t[0], t[1], gw, gh);
            return 0;
        }
        // "<file>|NRD_SIGNAL=<DIFF|SPEC|BOTH>|NRD_MODE=RADIANCE<suffix>": a single-lobe permutation binds only its own lobe's textures
        // (REBLUR_*.resources.hlsli), so the compact list is spread over the full slot table: C = always bound, D / S = lobe-only
        int signal = 0;
        const char* sigNames[4] = {nullptr, "|NRD_SIGNAL=DIFF|NRD_MODE=RADIANCE", "|NRD_SIGNAL=SPEC|NRD_MODE=RADIANCE", "|NRD_SIGNAL=BOTH|NRD_MODE=RADIANCE"};
        std::string file, suffix;
        for (int k = 1; k <= 3 && !signal; k++) {
            size_t at = id.find(sigNames[k]);
            if (at != std::string::npos) {
                signal = k;
                file = id.substr(0, at);
                suffix = id.substr(at + strlen(sigNames[k]));
            }
        }
        if (!signal) return 1;
        static Tex absent;  // never dereferenced: every access to a missing lobe is behind hasDiff() / hasSpec()
        Tex* f[25];
        auto spread = [&](const char* slots) -> bool {
            uint32_t next = 0;
            size_t n = strlen(slots);
            for (size_t i = 0; i < n; i++) {
                const bool bound = slots[i] == 'C' || (slots[i] == 'D' && (signal & 1)) || (slots[i] == 'S' && (signal & 2));
                if (bound && next >= texturesNum) return false;
                f[i] = bound ? &t[next++] : &absent;
            }
            return next == texturesNum;
        };
        if (file == "REBLUR_HitDistReconstruction.cs.hlsl" && (suffix == "|MODE_5X5=0" || suffix == "|MODE_5X5=1")) {
            if (!spread("CCCDSDS")) return 2;
            reblurHitDistReconstruction(cb, *f[0], *f[1], *f[2], *f[3], *f[4], *f[5], *f[6], gw, gh, suffix.back() == '1' ? 2 : 1, signal);
            return 0;
        }
        if (file == "REBLUR_PrePass.cs.hlsl" && suffix.empty()) {
            if (!spread("CCCDSDSS")) return 2;
            reblurPrePass(cb, *f[0], *f[1], *f[2], *f[3], *f[4], *f[5], *f[6], *f[7], gw, gh, robust, signal);
            return 0;
        }
        if (file == "REBLUR_TemporalAccumulation.cs.hlsl" && suffix.empty()) {
            if (!spread("CCCCCCCCDSDSDSDSSSCDSDSSC")) return 2;
            TaTextures a = {f[0], f[1], f[2], f[3], f[4], f[5], f[6], f[7], f[8], f[9], f[10], f[11], f[12], f[13], f[14], f[15], f[16], f[17],
                            f[18], f[19], f[20], f[21], f[22], f[23], f[24]};
            reblurTemporalAccumulation(cb, a, gw, gh, signal);
            return 0;
        }
        if (file == "REBLUR_HistoryFix.cs.hlsl" && suffix.empty()) {
            if (!spread("CCCCDSDSSDSDS")) return 2;
            HfTextures a = {f[0], f[1], f[2], f[3], f[4], f[5], f[6], f[7], f[8], f[9], f[10], f[11], f[12]};
            reblurHistoryFix(cb, a, gw, gh, quads, signal);
            return 0;
        }
        if (file == "REBLUR_Blur.cs.hlsl" && suffix.empty()) {
            if (!spread("CCCCDSCDS")) return 2;
            reblurBlur(cb, *f[0], *f[1], *f[2], *f[3], *f[4], *f[5], *f[6], *f[7], *f[8], gw, gh, quads, robust, signal);
            return 0;
        }
        if (file == "REBLUR_PostBlur.cs.hlsl" && suffix == "|TEMPORAL_STABILIZATION=1") {
            if (!spread("CCCCDSCDS")) return 2;
            reblurPostBlur(cb, *f[0], *f[1], *f[2], *f[3], *f[4], *f[5], *f[6], *f[7], *f[8], nullptr, nullptr, nullptr, true, gw, gh, quads, robust, signal);
            return 0;
        }
        if (file == "REBLUR_PostBlur.cs.hlsl" && suffix == "|TEMPORAL_STABILIZATION=0") {
            if (!spread("CCCCDSCDSCDS")) return 2;
            reblurPostBlur(cb, *f[0], *f[1], *f[2], *f[3], *f[4], *f[5], *f[6], *f[7], *f[8], f[9], f[10], f[11], false, gw, gh, quads, robust, signal);
            return 0;
        }
        if (file == "REBLUR_SplitScreen.cs.hlsl" && suffix.empty()) {
            if (!spread("CDSDS")) return 2;
            reblurSplitScreen(cb, *f[0], *f[1], *f[2], *f[3], *f[4], gw, gh, signal);
            return 0;
        }
        if (file == "REBLUR_TemporalStabilization.cs.hlsl" && suffix.empty()) {
            if (!spread("CCCCCSDSDSCCDSDS")) return 2;
            TsTextures a = {f[0], f[1], f[2], f[3], f[4], f[5], f[6], f[7], f[8], f[9], f[10], f[11], f[12], f[13], f[14], f[15]};
            reblurTemporalStabilization(cb, a, gw, gh, signal);
            return 0;
        }
        return 1;
    }
    if (startsWith(id, "SIGMA_")) return sigmaDispatch(id, constants, constantsSize, t, texturesNum, gw, gh);
    if (startsWith(id, "RELAX_")) return relaxDispatch(id, constants, constantsSize, t, texturesNum, gw, gh);
    if (startsWith(id, "REFERENCE_")) return referenceDispatch(id, constants, constantsSize, t, texturesNum, gw, gh);
    return 1;
}

__attribute__((visibility("default"))) void nrd_oracle_set_threads(int n) { omp_set_num_threads(n > 0 ? n : 1); }
__attribute__((visibility("default"))) int nrd_oracle_max_threads() { return omp_get_max_threads(); }

// Spot-check hooks for tests/test_oracle_math.py (compared against MathLib compiled from the reference tree)
__attribute__((visibility("default"))) uint16_t nrd_oracle_f32tof16(float f) { return f32tof16(f); }
__attribute__((visibility("default"))) float nrd_oracle_f16tof32(uint16_t h) { return f16tof32(h); }
__attribute__((visibility("default"))) uint32_t nrd_oracle_hash(uint32_t x) { return Sequence::Hash(x); }
__attribute__((visibility("default"))) float nrd_oracle_rng_first(uint32_t x, uint32_t y, uint32_t frame) {
    RngHash r;
    r.Initialize(x, y, frame);
    return r.GetFloat();
}
__attribute__((visibility("default"))) void nrd_oracle_pack_normal_roughness(const float* n, float roughness, float materialID, float* out4) {
    float4 p = NRD_FrontEnd_PackNormalAndRoughness(float3(n[0], n[1], n[2]), roughness, materialID);
    out4[0] = p.x; out4[1] = p.y; out4[2] = p.z; out4[3] = p.w;
}
__attribute__((visibility("default"))) void nrd_oracle_unpack_normal_roughness(const float* p4, float* out5) {
    float m;
    float4 r = NRD_FrontEnd_UnpackNormalAndRoughness(float4(p4[0], p4[1], p4[2], p4[3]), m);
    out5[0] = r.x; out5[1] = r.y; out5[2] = r.z; out5[3] = r.w; out5[4] = m;
}
__attribute__((visibility("default"))) float nrd_oracle_hitdist_normalization(float viewZ, float A, float B, float C, float roughness) {
    return _REBLUR_GetHitDistanceNormalization(viewZ, float3(A, B, C), roughness);
}
}
