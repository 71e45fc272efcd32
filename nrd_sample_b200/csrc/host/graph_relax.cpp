// RELAX_DIFFUSE_SPECULAR_SH / RELAX_DIFFUSE_SPECULAR pass graphs and per-frame constants.
// Pool layout and bindings: External/NRD/Source/Denoisers/Relax_DiffuseSpecularSh.hpp:13-382, Relax_DiffuseSpecular.hpp:13-312.
// Per-frame pass selection: External/NRD/Source/Relax.cpp:186-296. Constants: Relax.cpp:52-184.
#include <algorithm>
#include <cmath>

#include "pass_graph.h"

namespace nrdb {

namespace {

enum P : uint16_t {
    P_SPEC_ILLUM_PREV,
    P_SPEC_ILLUM_PREV_SH1,
    P_SPEC_ILLUM_RESPONSIVE_PREV,
    P_SPEC_ILLUM_RESPONSIVE_PREV_SH1,
    P_DIFF_ILLUM_PREV,
    P_DIFF_ILLUM_PREV_SH1,
    P_DIFF_ILLUM_RESPONSIVE_PREV,
    P_DIFF_ILLUM_RESPONSIVE_PREV_SH1,
    P_REFLECTION_HIT_T_CURR,
    P_REFLECTION_HIT_T_PREV,
    P_HISTORY_LENGTH_PREV,
    P_NORMAL_ROUGHNESS_PREV,
    P_MATERIAL_ID_PREV,
    P_VIEWZ_PREV,
};

enum T : uint16_t {
    T_SPEC_ILLUM_PING,
    T_SPEC_ILLUM_PING_SH1,
    T_SPEC_ILLUM_PONG,
    T_SPEC_ILLUM_PONG_SH1,
    T_DIFF_ILLUM_PING,
    T_DIFF_ILLUM_PING_SH1,
    T_DIFF_ILLUM_PONG,
    T_DIFF_ILLUM_PONG_SH1,
    T_SPEC_REPROJECTION_CONFIDENCE,
    T_TILES,
    T_HISTORY_LENGTH,
};

// Emission order; updateRelax() does arithmetic on these (Relax.cpp:187-199)
enum PassIndex : uint32_t {
    PASS_CLASSIFY_TILES = 0,
    PASS_HITDIST_RECONSTRUCTION = 1,  // 2: bit0 = 5x5
    PASS_PREPASS = 3,                 // 2: bit0 = after reconstruction
    PASS_TEMPORAL_ACCUMULATION = 5,   // 4: bit0 = confidence inputs, bit1 = disocclusion threshold mix
    PASS_HISTORY_FIX = 9,
    PASS_HISTORY_CLAMPING = 10,
    PASS_COPY = 11,
    PASS_ANTI_FIREFLY = 12,
    PASS_ATROUS = 13,                 // 2 x 5 binding variants: smem, even, odd, even-last, odd-last
    PASS_SPLIT_SCREEN = 23,
    PASS_VALIDATION = 24,
};
constexpr uint32_t kAtrousVariants = 5;
constexpr uint32_t kMaxAtrousPassNum = 8;

const uint32_t kCb = sizeof(RelaxConstants);

}  // namespace

// One graph for the six RELAX denoisers: RELAX_DIFFUSE_SPECULAR( _SH ) (Relax_DiffuseSpecular( Sh ).hpp), RELAX_DIFFUSE( _SH ) (Relax_Diffuse( Sh ).hpp) and
// RELAX_SPECULAR( _SH ) (Relax_Specular( Sh ).hpp). The single-lobe ones are the two-lobe graph without the other lobe's textures and bindings; the
// RADIANCE ones are the SH graph without the SH1 textures. The pool enums above name the textures, `Pm` / `Tr` map them to the pool order of the
// denoiser at hand.
void Graph::buildRelax(DenoiserState& d, bool diff, bool spec, bool sh) {
    new (&d.settings.relax) RelaxSettings();
    d.settingsSize = sizeof(RelaxSettings);

    uint16_t permIndex[16], tranIndex[16];
    for (int i = 0; i < 16; i++) permIndex[i] = tranIndex[i] = 0xFFFF;
    uint16_t nPerm = 0, nTran = 0;  // denoiser-local pool indices, in allocation order
    auto perm = [&](uint16_t which, Format f) { addPermanent(f); permIndex[which] = nPerm++; };
    auto tran = [&](uint16_t which, Format f, uint16_t downsample = 1) { addTransient(f, downsample); tranIndex[which] = nTran++; };
    const Format kIllum = Format::RGBA16_SFLOAT;
    if (diff && spec && !sh) {  // Relax_DiffuseSpecular.hpp:17-22: lobes interleaved
        perm(P_SPEC_ILLUM_PREV, kIllum);
        perm(P_DIFF_ILLUM_PREV, kIllum);
        perm(P_SPEC_ILLUM_RESPONSIVE_PREV, kIllum);
        perm(P_DIFF_ILLUM_RESPONSIVE_PREV, kIllum);
    } else {                    // lobe-major: { normal, ( SH1 ), responsive, ( SH1 ) } per lobe, specular first
        for (int lobe = 0; lobe < 2; lobe++) {
            if (!(lobe == 0 ? spec : diff)) continue;
            const uint16_t base = lobe == 0 ? P_SPEC_ILLUM_PREV : P_DIFF_ILLUM_PREV;
            perm(base + 0, kIllum);
            if (sh) perm(base + 1, kIllum);
            perm(base + 2, kIllum);
            if (sh) perm(base + 3, kIllum);
        }
    }
    if (spec) {
        perm(P_REFLECTION_HIT_T_CURR, Format::R16_SFLOAT);
        perm(P_REFLECTION_HIT_T_PREV, Format::R16_SFLOAT);
    }
    perm(P_HISTORY_LENGTH_PREV, Format::R8_UNORM);
    perm(P_NORMAL_ROUGHNESS_PREV, Format::RGBA8_UNORM);
    perm(P_MATERIAL_ID_PREV, Format::R8_UNORM);
    perm(P_VIEWZ_PREV, Format::R32_SFLOAT);

    for (int lobe = 0; lobe < 2; lobe++) {  // { ping, ( SH1 ), pong, ( SH1 ) } per lobe, specular first
        if (!(lobe == 0 ? spec : diff)) continue;
        const uint16_t base = lobe == 0 ? T_SPEC_ILLUM_PING : T_DIFF_ILLUM_PING;
        tran(base + 0, kIllum);
        if (sh) tran(base + 1, kIllum);
        tran(base + 2, kIllum);
        if (sh) tran(base + 3, kIllum);
    }
    if (spec) tran(T_SPEC_REPROJECTION_CONFIDENCE, Format::R8_UNORM);
    tran(T_TILES, Format::R8_UNORM, 16);
    tran(T_HISTORY_LENGTH, Format::R8_UNORM);

    auto U = [sh](ResourceType t) {
        if (!sh) {
            if (t == ResourceType::IN_SPEC_SH0) t = ResourceType::IN_SPEC_RADIANCE_HITDIST;
            else if (t == ResourceType::IN_DIFF_SH0) t = ResourceType::IN_DIFF_RADIANCE_HITDIST;
            else if (t == ResourceType::OUT_SPEC_SH0) t = ResourceType::OUT_SPEC_RADIANCE_HITDIST;
            else if (t == ResourceType::OUT_DIFF_SH0) t = ResourceType::OUT_DIFF_RADIANCE_HITDIST;
        }
        return Slot::user(t);
    };
    auto Pm = [&](uint16_t i) { return Slot::perm(permIndex[i]); };
    auto Tr = [&](uint16_t i) { return Slot::tran(tranIndex[i]); };
    const Slot dummy = U(ResourceType::IN_VIEWZ);
    const std::string signal = std::string("|NRD_SIGNAL=") + (diff && spec ? "BOTH" : (diff ? "DIFF" : "SPEC"));
    const std::string sig = signal + (sh ? "|NRD_MODE=SH" : "|NRD_MODE=RADIANCE");
    const std::string prefix = std::string("RELAX_") + (diff && spec ? "DiffuseSpecular" : (diff ? "Diffuse" : "Specular")) + (sh ? "Sh - " : " - ");
    auto name = [&](const char* pass) { return intern(prefix + pass); };
    // a lobe's binding exists only when the denoiser has that lobe
    auto inS = [&](Slot s, Slot swap = Slot()) { if (spec) in(s, swap); };
    auto inD = [&](Slot s, Slot swap = Slot()) { if (diff) in(s, swap); };
    auto outS = [&](Slot s, Slot swap = Slot()) { if (spec) out(s, swap); };
    auto outD = [&](Slot s, Slot swap = Slot()) { if (diff) out(s, swap); };

    beginPass(name("Classify tiles"));
    in(U(ResourceType::IN_VIEWZ));
    out(Tr(T_TILES));
    emit("RELAX_ClassifyTiles.cs.hlsl", 16, 16, kCb);

    for (int i = 0; i < 2; i++) {
        beginPass(name("Hit distance reconstruction"));
        in(Tr(T_TILES));
        in(U(ResourceType::IN_NORMAL_ROUGHNESS));
        in(U(ResourceType::IN_VIEWZ));
        inS(U(ResourceType::IN_SPEC_SH0));
        inD(U(ResourceType::IN_DIFF_SH0));
        outS(Tr(T_SPEC_ILLUM_PING));
        outD(Tr(T_DIFF_ILLUM_PING));
        emit("RELAX_HitDistReconstruction.cs.hlsl" + signal + "|NRD_MODE=RADIANCE" + (i ? "|MODE_5X5=1" : "|MODE_5X5=0"), 8, 8, kCb);
    }

    for (int i = 0; i < 2; i++) {
        const bool afterReconstruction = i & 1;
        beginPass(name("Pre-pass"));
        in(Tr(T_TILES));
        in(U(ResourceType::IN_NORMAL_ROUGHNESS));
        in(U(ResourceType::IN_VIEWZ));
        inS(afterReconstruction ? Tr(T_SPEC_ILLUM_PING) : U(ResourceType::IN_SPEC_SH0));
        inD(afterReconstruction ? Tr(T_DIFF_ILLUM_PING) : U(ResourceType::IN_DIFF_SH0));
        if (sh) inS(U(ResourceType::IN_SPEC_SH1));
        if (sh) inD(U(ResourceType::IN_DIFF_SH1));
        outS(U(ResourceType::OUT_SPEC_SH0));
        outD(U(ResourceType::OUT_DIFF_SH0));
        if (sh) outS(U(ResourceType::OUT_SPEC_SH1));
        if (sh) outD(U(ResourceType::OUT_DIFF_SH1));
        emit("RELAX_PrePass.cs.hlsl" + sig, 16, 16, kCb);
    }

    for (int i = 0; i < 4; i++) {
        const bool hasMix = (i >> 1) & 1, hasConfidence = i & 1;
        beginPass(name("Temporal accumulation"));
        in(Tr(T_TILES));
        in(U(ResourceType::IN_MV));
        in(U(ResourceType::IN_NORMAL_ROUGHNESS));
        in(U(ResourceType::IN_VIEWZ));
        in(hasMix ? U(ResourceType::IN_DISOCCLUSION_THRESHOLD_MIX) : dummy);
        in(Pm(P_NORMAL_ROUGHNESS_PREV));
        in(Pm(P_VIEWZ_PREV));
        in(Pm(P_HISTORY_LENGTH_PREV));
        in(Pm(P_MATERIAL_ID_PREV));
        inS(U(ResourceType::OUT_SPEC_SH0));
        inD(U(ResourceType::OUT_DIFF_SH0));
        inS(Pm(P_SPEC_ILLUM_RESPONSIVE_PREV));
        inD(Pm(P_DIFF_ILLUM_RESPONSIVE_PREV));
        inS(Pm(P_SPEC_ILLUM_PREV));
        inD(Pm(P_DIFF_ILLUM_PREV));
        inS(Pm(P_REFLECTION_HIT_T_PREV), Pm(P_REFLECTION_HIT_T_CURR));
        inS(hasConfidence ? U(ResourceType::IN_SPEC_CONFIDENCE) : dummy);
        inD(hasConfidence ? U(ResourceType::IN_DIFF_CONFIDENCE) : dummy);
        if (sh) inS(U(ResourceType::OUT_SPEC_SH1));
        if (sh) inD(U(ResourceType::OUT_DIFF_SH1));
        if (sh) inS(Pm(P_SPEC_ILLUM_RESPONSIVE_PREV_SH1));
        if (sh) inD(Pm(P_DIFF_ILLUM_RESPONSIVE_PREV_SH1));
        if (sh) inS(Pm(P_SPEC_ILLUM_PREV_SH1));
        if (sh) inD(Pm(P_DIFF_ILLUM_PREV_SH1));
        out(Tr(T_HISTORY_LENGTH));
        outS(Tr(T_SPEC_ILLUM_PING));
        outD(Tr(T_DIFF_ILLUM_PING));
        outS(Tr(T_SPEC_ILLUM_PONG));
        outD(Tr(T_DIFF_ILLUM_PONG));
        outS(Pm(P_REFLECTION_HIT_T_CURR), Pm(P_REFLECTION_HIT_T_PREV));
        outS(Tr(T_SPEC_REPROJECTION_CONFIDENCE));
        if (sh) outS(Tr(T_SPEC_ILLUM_PING_SH1));
        if (sh) outD(Tr(T_DIFF_ILLUM_PING_SH1));
        if (sh) outS(Tr(T_SPEC_ILLUM_PONG_SH1));
        if (sh) outD(Tr(T_DIFF_ILLUM_PONG_SH1));
        emit("RELAX_TemporalAccumulation.cs.hlsl" + sig, 8, 16, kCb);
    }

    beginPass(name("History fix"));
    in(Tr(T_TILES));
    in(Tr(T_HISTORY_LENGTH));
    in(U(ResourceType::IN_NORMAL_ROUGHNESS));
    in(U(ResourceType::IN_VIEWZ));
    inS(Tr(T_SPEC_ILLUM_PING));  // normal history
    inD(Tr(T_DIFF_ILLUM_PING));
    if (sh) inS(Tr(T_SPEC_ILLUM_PING_SH1));
    if (sh) inD(Tr(T_DIFF_ILLUM_PING_SH1));
    outS(Tr(T_SPEC_ILLUM_PONG));  // responsive history
    outD(Tr(T_DIFF_ILLUM_PONG));
    if (sh) outS(Tr(T_SPEC_ILLUM_PONG_SH1));
    if (sh) outD(Tr(T_DIFF_ILLUM_PONG_SH1));
    emit("RELAX_HistoryFix.cs.hlsl" + sig, 8, 8, kCb);

    beginPass(name("History clamping"));
    in(Tr(T_TILES));
    in(U(ResourceType::IN_VIEWZ));
    in(Tr(T_HISTORY_LENGTH));
    inS(U(ResourceType::OUT_SPEC_SH0));  // noisy input with the pre-blur applied
    inD(U(ResourceType::OUT_DIFF_SH0));
    inS(Tr(T_SPEC_ILLUM_PING));
    inD(Tr(T_DIFF_ILLUM_PING));
    inS(Tr(T_SPEC_ILLUM_PONG));
    inD(Tr(T_DIFF_ILLUM_PONG));
    if (sh) inS(Tr(T_SPEC_ILLUM_PING_SH1));
    if (sh) inD(Tr(T_DIFF_ILLUM_PING_SH1));
    if (sh) inS(Tr(T_SPEC_ILLUM_PONG_SH1));
    if (sh) inD(Tr(T_DIFF_ILLUM_PONG_SH1));
    out(Pm(P_HISTORY_LENGTH_PREV));
    outS(Pm(P_SPEC_ILLUM_PREV));
    outD(Pm(P_DIFF_ILLUM_PREV));
    outS(Pm(P_SPEC_ILLUM_RESPONSIVE_PREV));
    outD(Pm(P_DIFF_ILLUM_RESPONSIVE_PREV));
    if (sh) outS(Pm(P_SPEC_ILLUM_PREV_SH1));
    if (sh) outD(Pm(P_DIFF_ILLUM_PREV_SH1));
    if (sh) outS(Pm(P_SPEC_ILLUM_RESPONSIVE_PREV_SH1));
    if (sh) outD(Pm(P_DIFF_ILLUM_RESPONSIVE_PREV_SH1));
    emit("RELAX_HistoryClamping.cs.hlsl" + sig, 8, 8, kCb);

    beginPass(name("Copy"));
    inS(Pm(P_SPEC_ILLUM_PREV));
    inD(Pm(P_DIFF_ILLUM_PREV));
    outS(U(ResourceType::OUT_SPEC_SH0));
    outD(U(ResourceType::OUT_DIFF_SH0));
    emit("RELAX_Copy.cs.hlsl" + sig, 8, 8, kCb);

    beginPass(name("Anti-firefly"));
    in(Tr(T_TILES));
    in(U(ResourceType::IN_NORMAL_ROUGHNESS));
    in(U(ResourceType::IN_VIEWZ));
    inS(U(ResourceType::OUT_SPEC_SH0));
    inD(U(ResourceType::OUT_DIFF_SH0));
    outS(Pm(P_SPEC_ILLUM_PREV));
    outD(Pm(P_DIFF_ILLUM_PREV));
    emit("RELAX_AntiFirefly.cs.hlsl" + sig, 8, 8, kCb);

    for (int i = 0; i < 2; i++) {
        const bool hasConfidence = i & 1;
        for (uint32_t j = 0; j < kAtrousVariants; j++) {
            const bool smem = j == 0, even = j % 2 == 0, last = j > 2;
            beginPass(smem ? name("A-trous (SMEM)") : name("A-trous"));
            in(Tr(T_TILES));
            in(Tr(T_HISTORY_LENGTH));
            in(U(ResourceType::IN_NORMAL_ROUGHNESS));
            in(U(ResourceType::IN_VIEWZ));
            if (smem) {
                inS(Pm(P_SPEC_ILLUM_PREV));
                inD(Pm(P_DIFF_ILLUM_PREV));
            } else {
                inS(even ? Tr(T_SPEC_ILLUM_PONG) : Tr(T_SPEC_ILLUM_PING));
                inD(even ? Tr(T_DIFF_ILLUM_PONG) : Tr(T_DIFF_ILLUM_PING));
            }
            inS(Tr(T_SPEC_REPROJECTION_CONFIDENCE));
            inS(hasConfidence ? U(ResourceType::IN_SPEC_CONFIDENCE) : dummy);
            inD(hasConfidence ? U(ResourceType::IN_DIFF_CONFIDENCE) : dummy);
            if (smem) {
                if (sh) inS(Pm(P_SPEC_ILLUM_PREV_SH1));
                if (sh) inD(Pm(P_DIFF_ILLUM_PREV_SH1));
            } else {
                if (sh) inS(even ? Tr(T_SPEC_ILLUM_PONG_SH1) : Tr(T_SPEC_ILLUM_PING_SH1));
                if (sh) inD(even ? Tr(T_DIFF_ILLUM_PONG_SH1) : Tr(T_DIFF_ILLUM_PING_SH1));
            }
            if (last) {
                outS(U(ResourceType::OUT_SPEC_SH0));
                outD(U(ResourceType::OUT_DIFF_SH0));
            } else {
                outS(even ? Tr(T_SPEC_ILLUM_PING) : Tr(T_SPEC_ILLUM_PONG));
                outD(even ? Tr(T_DIFF_ILLUM_PING) : Tr(T_DIFF_ILLUM_PONG));
            }
            if (smem) {
                out(Pm(P_NORMAL_ROUGHNESS_PREV));
                out(Pm(P_MATERIAL_ID_PREV));
                out(Pm(P_VIEWZ_PREV));
            }
            if (last) {
                if (sh) outS(U(ResourceType::OUT_SPEC_SH1));
                if (sh) outD(U(ResourceType::OUT_DIFF_SH1));
            } else {
                if (sh) outS(even ? Tr(T_SPEC_ILLUM_PING_SH1) : Tr(T_SPEC_ILLUM_PONG_SH1));
                if (sh) outD(even ? Tr(T_DIFF_ILLUM_PING_SH1) : Tr(T_DIFF_ILLUM_PONG_SH1));
            }
            if (smem)
                emit("RELAX_AtrousSmem.cs.hlsl" + sig, 8, 8, kCb);
            else
                emit("RELAX_Atrous.cs.hlsl" + sig, 16, 16, kCb, 1, (kMaxAtrousPassNum - 2 + 1) / 2);
        }
    }

    beginPass(name("Split screen"));
    in(U(ResourceType::IN_VIEWZ));
    inD(U(ResourceType::IN_DIFF_SH0));
    inS(U(ResourceType::IN_SPEC_SH0));
    if (sh) inD(U(ResourceType::IN_DIFF_SH1));
    if (sh) inS(U(ResourceType::IN_SPEC_SH1));
    outD(U(ResourceType::OUT_DIFF_SH0));
    outS(U(ResourceType::OUT_SPEC_SH0));
    if (sh) outD(U(ResourceType::OUT_DIFF_SH1));
    if (sh) outS(U(ResourceType::OUT_SPEC_SH1));
    emit("RELAX_SplitScreen.cs.hlsl" + sig, 8, 16, kCb);

    beginPass(name("Validation"));
    in(U(ResourceType::IN_NORMAL_ROUGHNESS));
    in(U(ResourceType::IN_VIEWZ));
    in(U(ResourceType::IN_MV));
    in(Tr(T_HISTORY_LENGTH));
    out(U(ResourceType::OUT_VALIDATION));
    emit("RELAX_Validation.cs.hlsl", 8, 16, kCb, GRID_FROM_RESOURCE, 1);
}

void Graph::updateRelax(const DenoiserState& d) {
    const RelaxSettings& s = d.settings.relax;
    const bool reconstruct = s.hitDistanceReconstructionMode != HitDistanceReconstructionMode::OFF && s.checkerboardMode == CheckerboardMode::OFF;
    const uint32_t iterationNum = std::min(std::max(s.atrousIterationNum, 2u), kMaxAtrousPassNum);
    auto push = [&](uint32_t pass) { return fillRelaxConstants(s, pushDispatch(d, pass)); };

    if (m_common.splitScreen >= 1.0f) {
        push(PASS_SPLIT_SCREEN);
        return;
    }
    push(PASS_CLASSIFY_TILES);
    if (reconstruct) push(PASS_HITDIST_RECONSTRUCTION + (s.hitDistanceReconstructionMode == HitDistanceReconstructionMode::AREA_5X5 ? 1 : 0));
    push(PASS_PREPASS + (reconstruct ? 1 : 0));
    push(PASS_TEMPORAL_ACCUMULATION + (m_common.isDisocclusionThresholdMixAvailable ? 2 : 0) + (m_common.isHistoryConfidenceAvailable ? 1 : 0));
    push(PASS_HISTORY_FIX);
    push(PASS_HISTORY_CLAMPING);
    if (s.enableAntiFirefly) {
        push(PASS_COPY);
        push(PASS_ANTI_FIREFLY);
    }
    for (uint32_t i = 0; i < iterationNum; i++) {
        uint32_t pass = PASS_ATROUS + (m_common.isHistoryConfidenceAvailable ? kAtrousVariants : 0);
        if (i != 0) pass += 2 - (i & 1);
        if (i == iterationNum - 1) pass += 2;
        RelaxConstants* k = (RelaxConstants*)push(pass);
        if (k) {
            k->stepSize = 1u << i;
            k->isLastPass = i == iterationNum - 1 ? 1 : 0;
        }
    }
    if (m_common.splitScreen > 0.0f) push(PASS_SPLIT_SCREEN);
    if (m_common.enableValidation) push(PASS_VALIDATION);
}

void* Graph::fillRelaxConstants(const RelaxSettings& s, void* dst) {
    if (!dst) return dst;
    const CommonSettings& c = m_common;
    const FrameState& f = m_frame;
    const float resW = c.resourceSize[0], resH = c.resourceSize[1], resWp = c.resourceSizePrev[0], resHp = c.resourceSizePrev[1];
    const float rectW = c.rectSize[0], rectH = c.rectSize[1];

    // frustum basis in world space (Relax.cpp:44-74): right / up from the rows of worldToView, forward through the frustum centre
    auto basis = [](const Mat4& viewToClip, const Mat4& worldToView, const Mat4& viewToWorld, const float* frustum, float* right, float* up, float* fwd) {
        const float tanHalfFov = 1.0f / viewToClip.m[0];
        const float aspect = viewToClip.m[0] / viewToClip.m[5];
        for (int i = 0; i < 3; i++) {
            right[i] = worldToView.m[i * 4 + 0] * tanHalfFov;
            up[i] = worldToView.m[i * 4 + 1] * tanHalfFov * aspect;
        }
        const float v[4] = {0.5f * frustum[2] + frustum[0], 0.5f * frustum[3] + frustum[1], 1.0f, 0.0f};
        for (int i = 0; i < 3; i++) {
            float r = v[0] * viewToWorld.m[0 + i];
            r = v[1] * viewToWorld.m[4 + i] + r;
            r = v[2] * viewToWorld.m[8 + i] + r;
            r = v[3] * viewToWorld.m[12 + i] + r;
            fwd[i] = r;
        }
        right[3] = up[3] = fwd[3] = 0.0f;
    };

    const bool reset = c.accumulationMode != AccumulationMode::CONTINUE;
    const float thresholdBonus = (1.0f + f.jitterDelta) / rectH;
    uint32_t diffCheckerboard = 2, specCheckerboard = 2;
    if (s.checkerboardMode == CheckerboardMode::BLACK) { diffCheckerboard = 0; specCheckerboard = 1; }
    else if (s.checkerboardMode == CheckerboardMode::WHITE) { diffCheckerboard = 1; specCheckerboard = 0; }
    auto sat = [](float x) { return std::min(std::max(x, 0.0f), 1.0f); };
    auto frames = [&](uint32_t n) { return reset ? 0.0f : (float)std::min(n, RELAX_MAX_HISTORY_FRAME_NUM); };

    RelaxConstants& k = *(RelaxConstants*)dst;
    k.worldToClip = f.worldToClip;
    k.worldToClipPrev = f.worldToClipPrev;
    k.worldToViewPrev = f.worldToViewPrev;
    k.worldPrevToWorld = f.worldPrevToWorld;
    memcpy(k.rotatorPre, f.rotatorPre, 16);
    basis(f.viewToClip, f.worldToView, f.viewToWorld, f.frustum, k.frustumRight, k.frustumUp, k.frustumForward);
    basis(f.viewToClipPrev, f.worldToViewPrev, f.viewToWorldPrev, f.frustumPrev, k.prevFrustumRight, k.prevFrustumUp, k.prevFrustumForward);
    for (int i = 0; i < 3; i++) {
        k.cameraDelta[i] = f.cameraDelta[i];
        k.mvScale[i] = c.motionVectorScale[i];
    }
    k.mvScale[3] = c.isMotionVectorInWorldSpace ? 1.0f : 0.0f;
    k.jitter[0] = c.cameraJitter[0]; k.jitter[1] = c.cameraJitter[1];
    k.resolutionScale[0] = rectW / resW; k.resolutionScale[1] = rectH / resH;
    k.rectOffset[0] = float(c.rectOrigin[0]) / resW; k.rectOffset[1] = float(c.rectOrigin[1]) / resH;
    k.resourceSizeInv[0] = 1.0f / resW; k.resourceSizeInv[1] = 1.0f / resH;
    k.resourceSize[0] = resW; k.resourceSize[1] = resH;
    k.rectSizeInv[0] = 1.0f / rectW; k.rectSizeInv[1] = 1.0f / rectH;
    k.rectSizePrev[0] = c.rectSizePrev[0]; k.rectSizePrev[1] = c.rectSizePrev[1];
    k.resourceSizeInvPrev[0] = 1.0f / resWp; k.resourceSizeInvPrev[1] = 1.0f / resHp;
    k.printfAt[0] = c.printfAt[0]; k.printfAt[1] = c.printfAt[1];
    k.rectOrigin[0] = c.rectOrigin[0]; k.rectOrigin[1] = c.rectOrigin[1];
    k.rectSize[0] = c.rectSize[0]; k.rectSize[1] = c.rectSize[1];
    k.specMaxAccumulatedFrameNum = frames(s.specularMaxAccumulatedFrameNum);
    k.specMaxFastAccumulatedFrameNum = frames(s.specularMaxFastAccumulatedFrameNum);
    k.diffMaxAccumulatedFrameNum = frames(s.diffuseMaxAccumulatedFrameNum);
    k.diffMaxFastAccumulatedFrameNum = frames(s.diffuseMaxFastAccumulatedFrameNum);
    k.disocclusionThreshold = c.disocclusionThreshold + thresholdBonus;
    k.disocclusionThresholdAlternate = c.disocclusionThresholdAlternate + thresholdBonus;
    k.cameraAttachedReflectionMaterialID = c.cameraAttachedReflectionMaterialID;
    k.strandMaterialID = c.strandMaterialID;
    k.strandThickness = c.strandThickness;
    k.roughnessFraction = s.roughnessFraction;
    k.specVarianceBoost = s.specularVarianceBoost;
    k.splitScreen = c.splitScreen;
    k.diffBlurRadius = s.diffusePrepassBlurRadius;
    k.specBlurRadius = s.specularPrepassBlurRadius;
    k.depthThreshold = s.depthThreshold;
    k.lobeAngleFraction = s.lobeAngleFraction;
    k.specLobeAngleSlack = s.specularLobeAngleSlack * (3.14159265358979323846f / 180.0f);
    k.historyFixEdgeStoppingNormalPower = s.historyFixEdgeStoppingNormalPower;
    k.roughnessEdgeStoppingRelaxation = s.roughnessEdgeStoppingRelaxation;
    k.normalEdgeStoppingRelaxation = s.normalEdgeStoppingRelaxation;
    k.fastHistoryClampingSigmaScale = s.fastHistoryClampingSigmaScale;
    k.historyAccelerationAmount = s.antilagSettings.accelerationAmount;
    k.historyResetTemporalSigmaScale = s.antilagSettings.temporalSigmaScale;
    k.historyResetSpatialSigmaScale = s.antilagSettings.spatialSigmaScale;
    k.historyResetAmount = s.antilagSettings.resetAmount;
    k.denoisingRange = c.denoisingRange;
    k.specPhiLuminance = s.specularPhiLuminance;
    k.diffPhiLuminance = s.diffusePhiLuminance;
    k.diffMaxLuminanceRelativeDifference = -std::log(sat(s.diffuseMinLuminanceWeight));
    k.specMaxLuminanceRelativeDifference = -std::log(sat(s.specularMinLuminanceWeight));
    k.luminanceEdgeStoppingRelaxation = s.roughnessEdgeStoppingRelaxation;  // sic (Relax.cpp:153)
    k.confidenceDrivenRelaxationMultiplier = s.confidenceDrivenRelaxationMultiplier;
    k.confidenceDrivenLuminanceEdgeStoppingRelaxation = s.confidenceDrivenLuminanceEdgeStoppingRelaxation;
    k.confidenceDrivenNormalEdgeStoppingRelaxation = s.confidenceDrivenNormalEdgeStoppingRelaxation;
    k.debug = c.debug;
    k.orthoMode = f.orthoMode;
    k.unproject = 1.0f / (0.5f * rectH * f.projectY);
    k.framerateScale = std::min(std::max(16.66f / f.timeDelta, 0.25f), 4.0f);
    k.checkerboardResolveAccumSpeed = f.checkerboardResolveAccumSpeed;
    k.historyFixFrameNum = s.historyFixFrameNum + 1.0f;
    k.historyFixBasePixelStride = (float)s.historyFixBasePixelStride;
    k.historyFixAlternatePixelStride = (float)s.historyFixAlternatePixelStride;
    k.historyFixAlternatePixelStrideMaterialID = c.historyFixAlternatePixelStrideMaterialID;
    k.historyThreshold = (float)s.spatialVarianceEstimationHistoryThreshold;
    k.viewZScale = c.viewZScale;
    k.minHitDistanceWeight = s.minHitDistanceWeight * 2.0f;
    k.diffMinMaterial = s.minMaterialForDiffuse;
    k.specMinMaterial = s.minMaterialForSpecular;
    k.roughnessEdgeStoppingEnabled = s.enableRoughnessEdgeStopping ? 1 : 0;
    k.frameIndex = c.frameIndex;
    k.diffCheckerboard = diffCheckerboard;
    k.specCheckerboard = specCheckerboard;
    k.hasHistoryConfidence = c.isHistoryConfidenceAvailable ? 1 : 0;
    k.hasDisocclusionThresholdMix = c.isDisocclusionThresholdMixAvailable ? 1 : 0;
    k.resetHistory = reset ? 1 : 0;
    return dst;
}

}  // namespace nrdb
