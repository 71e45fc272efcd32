// shaderIdentifier -> PipelineKey ( see ../pipeline_key.h ). The table is the permutation list of the reference's Shaders.cfg:3-72: one row per shader file
// with the defines its permutations carry and the NRD_MODE values it is built for. Plain C string work, done once per pipeline at instance creation.
#include "../pipeline_key.h"

#include <cstring>

namespace nrdk {
namespace {

enum : uint32_t { D_SIGNAL = 1, D_MODE = 2, D_5X5 = 4, D_TS = 8, D_TRANSLUCENCY = 16, D_FIRST_PASS = 32, D_FLOAT = 64 };
enum : uint32_t { M_RADIANCE = 1, M_SH = 2, M_OCCLUSION = 4, M_DO = 8 };

struct Row {
    const char* file;
    PipelineFamily family;
    PipelinePass pass;
    uint32_t defines;   // exactly these keys must be present
    uint32_t modes;     // allowed NRD_MODE values ( when D_MODE )
};

const Row kRows[] = {
    {"Clear.cs.hlsl", FAMILY_CLEAR, PASS_NONE, D_FLOAT, 0},
    {"REBLUR_ClassifyTiles.cs.hlsl", FAMILY_REBLUR, REBLUR_CLASSIFY_TILES, 0, 0},
    {"REBLUR_HitDistReconstruction.cs.hlsl", FAMILY_REBLUR, REBLUR_HITDIST_RECONSTRUCTION, D_SIGNAL | D_MODE | D_5X5, M_RADIANCE | M_OCCLUSION},
    {"REBLUR_PrePass.cs.hlsl", FAMILY_REBLUR, REBLUR_PREPASS, D_SIGNAL | D_MODE, M_RADIANCE | M_SH | M_DO},
    {"REBLUR_TemporalAccumulation.cs.hlsl", FAMILY_REBLUR, REBLUR_TEMPORAL_ACCUMULATION, D_SIGNAL | D_MODE, M_RADIANCE | M_SH | M_OCCLUSION | M_DO},
    {"REBLUR_HistoryFix.cs.hlsl", FAMILY_REBLUR, REBLUR_HISTORY_FIX, D_SIGNAL | D_MODE, M_RADIANCE | M_SH | M_OCCLUSION | M_DO},
    {"REBLUR_Blur.cs.hlsl", FAMILY_REBLUR, REBLUR_BLUR, D_SIGNAL | D_MODE, M_RADIANCE | M_SH | M_OCCLUSION | M_DO},
    {"REBLUR_PostBlur.cs.hlsl", FAMILY_REBLUR, REBLUR_POST_BLUR, D_SIGNAL | D_MODE | D_TS, M_RADIANCE | M_SH | M_OCCLUSION | M_DO},
    {"REBLUR_TemporalStabilization.cs.hlsl", FAMILY_REBLUR, REBLUR_TEMPORAL_STABILIZATION, D_SIGNAL | D_MODE, M_RADIANCE | M_SH | M_DO},
    {"REBLUR_SplitScreen.cs.hlsl", FAMILY_REBLUR, REBLUR_SPLIT_SCREEN, D_SIGNAL | D_MODE, M_RADIANCE | M_SH},
    {"REBLUR_Validation.cs.hlsl", FAMILY_REBLUR, REBLUR_VALIDATION, 0, 0},
    {"RELAX_ClassifyTiles.cs.hlsl", FAMILY_RELAX, RELAX_CLASSIFY_TILES, 0, 0},
    {"RELAX_HitDistReconstruction.cs.hlsl", FAMILY_RELAX, RELAX_HITDIST_RECONSTRUCTION, D_SIGNAL | D_MODE | D_5X5, M_RADIANCE},
    {"RELAX_PrePass.cs.hlsl", FAMILY_RELAX, RELAX_PREPASS, D_SIGNAL | D_MODE, M_RADIANCE | M_SH},
    {"RELAX_TemporalAccumulation.cs.hlsl", FAMILY_RELAX, RELAX_TEMPORAL_ACCUMULATION, D_SIGNAL | D_MODE, M_RADIANCE | M_SH},
    {"RELAX_HistoryFix.cs.hlsl", FAMILY_RELAX, RELAX_HISTORY_FIX, D_SIGNAL | D_MODE, M_RADIANCE | M_SH},
    {"RELAX_HistoryClamping.cs.hlsl", FAMILY_RELAX, RELAX_HISTORY_CLAMPING, D_SIGNAL | D_MODE, M_RADIANCE | M_SH},
    {"RELAX_Copy.cs.hlsl", FAMILY_RELAX, RELAX_COPY, D_SIGNAL | D_MODE, M_RADIANCE | M_SH},
    {"RELAX_AntiFirefly.cs.hlsl", FAMILY_RELAX, RELAX_ANTI_FIREFLY, D_SIGNAL | D_MODE, M_RADIANCE | M_SH},
    {"RELAX_AtrousSmem.cs.hlsl", FAMILY_RELAX, RELAX_ATROUS_SMEM, D_SIGNAL | D_MODE, M_RADIANCE | M_SH},
    {"RELAX_Atrous.cs.hlsl", FAMILY_RELAX, RELAX_ATROUS, D_SIGNAL | D_MODE, M_RADIANCE | M_SH},
    {"RELAX_SplitScreen.cs.hlsl", FAMILY_RELAX, RELAX_SPLIT_SCREEN, D_SIGNAL | D_MODE, M_RADIANCE | M_SH},
    {"RELAX_Validation.cs.hlsl", FAMILY_RELAX, RELAX_VALIDATION, 0, 0},
    {"SIGMA_ClassifyTiles.cs.hlsl", FAMILY_SIGMA, SIGMA_CLASSIFY_TILES, D_TRANSLUCENCY, 0},
    {"SIGMA_SmoothTiles.cs.hlsl", FAMILY_SIGMA, SIGMA_SMOOTH_TILES, 0, 0},
    {"SIGMA_Copy.cs.hlsl", FAMILY_SIGMA, SIGMA_COPY, 0, 0},
    {"SIGMA_Blur.cs.hlsl", FAMILY_SIGMA, SIGMA_BLUR, D_TRANSLUCENCY | D_FIRST_PASS, 0},
    {"SIGMA_TemporalStabilization.cs.hlsl", FAMILY_SIGMA, SIGMA_TEMPORAL_STABILIZATION, D_TRANSLUCENCY, 0},
    {"SIGMA_SplitScreen.cs.hlsl", FAMILY_SIGMA, SIGMA_SPLIT_SCREEN, D_TRANSLUCENCY, 0},
    {"REFERENCE_TemporalAccumulation.cs.hlsl", FAMILY_REFERENCE, REFERENCE_TEMPORAL_ACCUMULATION, 0, 0},
    {"REFERENCE_Copy.cs.hlsl", FAMILY_REFERENCE, REFERENCE_COPY, 0, 0},
};

bool tokenIs(const char* token, size_t len, const char* text) { return strlen(text) == len && memcmp(token, text, len) == 0; }

// "0" / "1" -> out; anything else fails
bool flagValue(const char* v, size_t len, bool& out) {
    if (len != 1 || (v[0] != '0' && v[0] != '1')) return false;
    out = v[0] == '1';
    return true;
}

}  // namespace

PipelineKey resolvePipeline(const char* id) {
    PipelineKey key, unknown;
    unknown.id = key.id = id ? id : "";
    if (!id) return unknown;
    const char* bar = strchr(id, '|');
    const size_t fileLen = bar ? (size_t)(bar - id) : strlen(id);
    const Row* row = nullptr;
    for (const Row& r : kRows)
        if (tokenIs(id, fileLen, r.file)) row = &r;
    if (!row) return unknown;
    key.family = row->family;
    key.pass = row->pass;
    uint32_t seen = 0;
    bool ignoredFlag = false;
    for (const char* p = bar; p && *p;) {   // "|KEY=VALUE" ...
        const char* k = p + 1;
        const char* next = strchr(k, '|');
        const size_t len = next ? (size_t)(next - k) : strlen(k);
        const char* eq = (const char*)memchr(k, '=', len);
        if (!eq) return unknown;
        const size_t klen = (size_t)(eq - k), vlen = len - klen - 1;
        const char* v = eq + 1;
        uint32_t bit = 0;
        if (tokenIs(k, klen, "NRD_SIGNAL")) {
            bit = D_SIGNAL;
            key.signal = tokenIs(v, vlen, "DIFF") ? 1 : (tokenIs(v, vlen, "SPEC") ? 2 : (tokenIs(v, vlen, "BOTH") ? 3 : 0));
            if (!key.signal) return unknown;
        } else if (tokenIs(k, klen, "NRD_MODE")) {
            bit = D_MODE;
            int m = tokenIs(v, vlen, "RADIANCE") ? 0 : (tokenIs(v, vlen, "SH") ? 1 : (tokenIs(v, vlen, "OCCLUSION") ? 2 : (tokenIs(v, vlen, "DO") ? 3 : -1)));
            if (m < 0 || !(row->modes & (1u << m))) return unknown;
            key.mode = (uint8_t)m;
            key.hasMode = true;
        } else if (tokenIs(k, klen, "MODE_5X5")) {
            bit = D_5X5;
            if (!flagValue(v, vlen, key.mode5x5)) return unknown;
        } else if (tokenIs(k, klen, "TEMPORAL_STABILIZATION")) {
            bit = D_TS;
            if (!flagValue(v, vlen, key.temporalStabilization)) return unknown;
        } else if (tokenIs(k, klen, "TRANSLUCENCY")) {
            bit = D_TRANSLUCENCY;
            if (!flagValue(v, vlen, key.translucency)) return unknown;
        } else if (tokenIs(k, klen, "FIRST_PASS")) {
            bit = D_FIRST_PASS;
            if (!flagValue(v, vlen, key.firstPass)) return unknown;
        } else if (tokenIs(k, klen, "FLOAT")) {
            bit = D_FLOAT;
            if (!flagValue(v, vlen, ignoredFlag)) return unknown;   // the clear value is 0 either way: one kernel zeroes the bytes
        } else
            return unknown;
        if (seen & bit) return unknown;
        seen |= bit;
        p = next;
    }
    // a clear may come with or without its FLOAT define ( the executor's own callers pass the bare file name )
    if (row->family == FAMILY_CLEAR ? (seen & ~D_FLOAT) != 0 : seen != row->defines) return unknown;
    return key;
}

}  // namespace nrdk
