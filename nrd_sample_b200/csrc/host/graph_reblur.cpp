// REBLUR_DIFFUSE / REBLUR_SPECULAR / REBLUR_DIFFUSE_SPECULAR pass graphs and per-frame constants.
// Pool layout and bindings: External/NRD/Source/Denoisers/Reblur_Diffuse.hpp, Reblur_Specular.hpp, Reblur_DiffuseSpecular.hpp.
// Per-frame pass selection: External/NRD/Source/Reblur.cpp:98-199. Constants: Reblur.cpp:283-394.
#include <algorithm>
#include <cmath>

#include "pass_graph.h"

namespace nrdb {

namespace {

// Index of each pass (and its permutations) in emission order; updateReblur() does arithmetic on these
enum PassIndex : uint32_t {
    PASS_CLASSIFY_TILES = 0,
    PASS_HITDIST_RECONSTRUCTION = 1,  // 4 permutations: bit0 = prepass follows, bit1 = 5x5
    PASS_PREPASS = 5,                 // 2: bit0 = reads reconstruction output
    PASS_TEMPORAL_ACCUMULATION = 7,   // 8: bit0 = after prepass/reconstruction, bit1 = confidence inputs, bit2 = threshold mix
    PASS_HISTORY_FIX = 15,
    PASS_BLUR = 16,
    PASS_POST_BLUR = 17,              // 2: bit0 = temporal stabilization follows
    PASS_TEMPORAL_STABILIZATION = 19,
    PASS_SPLIT_SCREEN = 20,
    PASS_VALIDATION = 21,
};

const uint32_t kCb = sizeof(ReblurConstants);

// g_ReblurProps ( Reblur.cpp:85-96 )
bool hasDiffuse(Denoiser d) {
    return d == Denoiser::REBLUR_DIFFUSE || d == Denoiser::REBLUR_DIFFUSE_SPECULAR || d == Denoiser::REBLUR_DIFFUSE_SH || d == Denoiser::REBLUR_DIFFUSE_SPECULAR_SH ||
           d == Denoiser::REBLUR_DIFFUSE_OCCLUSION || d == Denoiser::REBLUR_DIFFUSE_SPECULAR_OCCLUSION || d == Denoiser::REBLUR_DIFFUSE_DIRECTIONAL_OCCLUSION;
}
bool hasSpecular(Denoiser d) {
    return d == Denoiser::REBLUR_SPECULAR || d == Denoiser::REBLUR_DIFFUSE_SPECULAR || d == Denoiser::REBLUR_SPECULAR_SH || d == Denoiser::REBLUR_DIFFUSE_SPECULAR_SH ||
           d == Denoiser::REBLUR_SPECULAR_OCCLUSION || d == Denoiser::REBLUR_DIFFUSE_SPECULAR_OCCLUSION;
}

// Pass indices of the occlusion denoisers ( Update_ReblurOcclusion, Reblur.cpp:204-215 ): no pre-pass, no stabilization
enum OcclusionPassIndex : uint32_t {
    OCC_CLASSIFY_TILES = 0,
    OCC_HITDIST_RECONSTRUCTION = 1,  // 2 permutations: bit0 = 5x5
    OCC_TEMPORAL_ACCUMULATION = 3,   // 8: bit0 = after reconstruction, bit1 = confidence inputs, bit2 = threshold mix
    OCC_HISTORY_FIX = 11,
    OCC_BLUR = 12,
    OCC_POST_BLUR = 13,
    OCC_SPLIT_SCREEN = 14,
    OCC_VALIDATION = 15,
};

}  // namespace

// One builder for REBLUR_DIFFUSE / REBLUR_SPECULAR / REBLUR_DIFFUSE_SPECULAR (NRD_MODE = RADIANCE): the reference's three .hpp files are the
// same graph with the bindings of the absent lobe removed (Reblur_Diffuse.hpp:15-249, Reblur_Specular.hpp:15-260, Reblur_DiffuseSpecular.hpp:17-297).
// Pool layout: permanent = prev viewZ / normal+roughness / internal data, then per lobe { history, fast history, stabilized luma ping / pong },
// then the specular hit-distance-for-tracking ping / pong; transient = data1 (RG8 for both lobes, R8 for one), data2 (R32_UINT, R8_UINT without
// specular), hit distance for tracking (specular), per lobe { tmp2, fast history }, tiles.
// sh: REBLUR_DIFFUSE_SH / REBLUR_SPECULAR_SH / REBLUR_DIFFUSE_SPECULAR_SH (NRD_MODE = SH; Reblur_DiffuseSh.hpp, Reblur_SpecularSh.hpp,
// Reblur_DiffuseSpecularSh.hpp): the same graph over IN / OUT_*_SH0 with a second RGBA16F per lobe ( SH1 ) carried through every pass — the
// user's OUT_*_SH1 as "SH_TEMP1", one more permanent ( SH history, after the lobe's stabilized pair ) and transient ( SH_TMP2, after the lobe's fast history ).
// directional: REBLUR_DIFFUSE_DIRECTIONAL_OCCLUSION ( NRD_MODE = DO, Reblur_DiffuseDirectionalOcclusion.hpp ): the diffuse-only graph over IN / OUT_DIFF_DIRECTION_HITDIST
// with 16-bit SNORM signal textures and an 8-bit UNORM fast history ( Reblur.cpp:38-39 ); reconstruction and split screen keep the RADIANCE permutation.
void Graph::buildReblur(DenoiserState& d, bool diff, bool spec, bool sh, bool directional) {
    new (&d.settings.reblur) ReblurSettings();
    d.settingsSize = sizeof(ReblurSettings);

    const Format kRadiance = directional ? Format::RGBA16_SNORM : Format::RGBA16_SFLOAT, kFast = directional ? Format::R8_UNORM : Format::R16_SFLOAT;
    uint16_t nPerm = 0, nTran = 0;
    auto perm = [&](Format f) { addPermanent(f); return nPerm++; };
    auto tran = [&](Format f, uint16_t ds = 1) { addTransient(f, ds); return nTran++; };
    const uint16_t P_PREV_VIEWZ = perm(Format::R32_SFLOAT);
    const uint16_t P_PREV_NORMAL_ROUGHNESS = perm(Format::R10_G10_B10_A2_UNORM);  // must match the IN_NORMAL_ROUGHNESS encoding
    const uint16_t P_PREV_INTERNAL_DATA = perm(Format::R16_UINT);                 // 6b diff frames | 6b spec frames | 4b material
    uint16_t P_DIFF_HISTORY = 0, P_DIFF_FAST_HISTORY = 0, P_DIFF_STABILIZED_PING = 0, P_DIFF_STABILIZED_PONG = 0;
    uint16_t P_DIFF_SH_HISTORY = 0, P_SPEC_SH_HISTORY = 0, T_DIFF_SH_TMP2 = 0, T_SPEC_SH_TMP2 = 0;
    uint16_t P_SPEC_HISTORY = 0, P_SPEC_FAST_HISTORY = 0, P_SPEC_STABILIZED_PING = 0, P_SPEC_STABILIZED_PONG = 0, P_SPEC_HITDIST_TRACKING_PING = 0, P_SPEC_HITDIST_TRACKING_PONG = 0;
    if (diff) {
        P_DIFF_HISTORY = perm(kRadiance);
        P_DIFF_FAST_HISTORY = perm(kFast);
        P_DIFF_STABILIZED_PING = perm(Format::R16_SFLOAT);
        P_DIFF_STABILIZED_PONG = perm(Format::R16_SFLOAT);
        if (sh) P_DIFF_SH_HISTORY = perm(kRadiance);
    }
    if (spec) {
        P_SPEC_HISTORY = perm(kRadiance);
        P_SPEC_FAST_HISTORY = perm(kFast);
        P_SPEC_STABILIZED_PING = perm(Format::R16_SFLOAT);
        P_SPEC_STABILIZED_PONG = perm(Format::R16_SFLOAT);
        if (sh) P_SPEC_SH_HISTORY = perm(kRadiance);
        P_SPEC_HITDIST_TRACKING_PING = perm(Format::R16_SFLOAT);
        P_SPEC_HITDIST_TRACKING_PONG = perm(Format::R16_SFLOAT);
    }

    const uint16_t T_DATA1 = tran(diff && spec ? Format::RG8_UNORM : Format::R8_UNORM);
    const uint16_t T_DATA2 = tran(spec ? Format::R32_UINT : Format::R8_UINT);
    uint16_t T_SPEC_HITDIST_TRACKING = 0, T_DIFF_TMP2 = 0, T_DIFF_FAST = 0, T_SPEC_TMP2 = 0, T_SPEC_FAST = 0;
    if (spec) T_SPEC_HITDIST_TRACKING = tran(Format::R16_SFLOAT);
    if (diff) {
        T_DIFF_TMP2 = tran(kRadiance);
        T_DIFF_FAST = tran(kFast);
        if (sh) T_DIFF_SH_TMP2 = tran(kRadiance);
    }
    if (spec) {
        T_SPEC_TMP2 = tran(kRadiance);
        T_SPEC_FAST = tran(kFast);
        if (sh) T_SPEC_SH_TMP2 = tran(kRadiance);
    }
    // Reblur_DiffuseSpecularSh.hpp:59-82 adds ELEVEN transient textures for its ten names: a third RGBA16F sits where the enum says TILES, and the
    // R8 texture at 1/16 resolution that follows is never bound. The passes therefore keep their sky-tile mask in the top-left corner of a
    // full-resolution RGBA16F texture. Kept as is: pool sizes and binding indices are part of the stream an integration sees.
    uint16_t T_TILES;
    if (diff && spec && sh) {
        T_TILES = tran(kRadiance);
        tran(Format::R8_UNORM, 16);
    } else {
        T_TILES = tran(Format::R8_UNORM, 16);
    }

    auto U = [](ResourceType t) { return Slot::user(t); };
    auto Pm = [](uint16_t i) { return Slot::perm(i); };
    auto Tr = [](uint16_t i) { return Slot::tran(i); };
    // The user's output textures double as scratch ("TEMP1") between passes
    const ResourceType inDiff = directional ? ResourceType::IN_DIFF_DIRECTION_HITDIST : (sh ? ResourceType::IN_DIFF_SH0 : ResourceType::IN_DIFF_RADIANCE_HITDIST), inSpec = sh ? ResourceType::IN_SPEC_SH0 : ResourceType::IN_SPEC_RADIANCE_HITDIST;
    const ResourceType outDiff = directional ? ResourceType::OUT_DIFF_DIRECTION_HITDIST : (sh ? ResourceType::OUT_DIFF_SH0 : ResourceType::OUT_DIFF_RADIANCE_HITDIST), outSpec = sh ? ResourceType::OUT_SPEC_SH0 : ResourceType::OUT_SPEC_RADIANCE_HITDIST;
    const Slot diffTemp1 = U(outDiff), specTemp1 = U(outSpec);
    const Slot diffTemp2 = Tr(T_DIFF_TMP2), specTemp2 = Tr(T_SPEC_TMP2);
    const Slot diffShTemp1 = U(ResourceType::OUT_DIFF_SH1), specShTemp1 = U(ResourceType::OUT_SPEC_SH1), diffShTemp2 = Tr(T_DIFF_SH_TMP2), specShTemp2 = Tr(T_SPEC_SH_TMP2);
    const Slot dummy = U(ResourceType::IN_VIEWZ);  // bound where an optional input is absent
    const std::string sigSignal = std::string("|NRD_SIGNAL=") + (diff && spec ? "BOTH" : (diff ? "DIFF" : "SPEC"));
    const std::string sig = sigSignal + (directional ? "|NRD_MODE=DO" : (sh ? "|NRD_MODE=SH" : "|NRD_MODE=RADIANCE"));
    // Reblur_DiffuseSh.hpp never defines DENOISER_NAME ( the other files do ), so its pass names carry the macro's own name
    const std::string prefix = directional ? std::string("REBLUR_DirectionalOcclusion - ") : (sh && diff && !spec) ? std::string("DENOISER_NAME - ")
                                                     : std::string("REBLUR_") + (diff && spec ? "DiffuseSpecular" : (diff ? "Diffuse" : "Specular")) + (sh ? "Sh - " : " - ");
    auto name = [&](const char* pass) { return intern(prefix + pass); };
    // bind helpers: a lobe's binding exists only when the denoiser has that lobe
    auto inD = [&](Slot s, Slot swap = Slot()) { if (diff) in(s, swap); };
    auto inS = [&](Slot s, Slot swap = Slot()) { if (spec) in(s, swap); };
    auto outD = [&](Slot s, Slot swap = Slot()) { if (diff) out(s, swap); };
    auto outS = [&](Slot s, Slot swap = Slot()) { if (spec) out(s, swap); };
    auto inShD = [&](Slot s) { if (diff && sh) in(s); };
    auto inShS = [&](Slot s) { if (spec && sh) in(s); };
    auto outShD = [&](Slot s) { if (diff && sh) out(s); };
    auto outShS = [&](Slot s) { if (spec && sh) out(s); };

    beginPass(name("Classify tiles"));
    in(U(ResourceType::IN_VIEWZ));
    out(Tr(T_TILES));
    emit("REBLUR_ClassifyTiles.cs.hlsl", 16, 16, kCb);

    for (int i = 0; i < 4; i++) {
        bool is5x5 = (i >> 1) & 1, prepassFollows = i & 1;
        beginPass(name("Hit distance reconstruction"));
        in(Tr(T_TILES));
        in(U(ResourceType::IN_NORMAL_ROUGHNESS));
        in(U(ResourceType::IN_VIEWZ));
        inD(U(inDiff));
        inS(U(inSpec));
        outD(prepassFollows ? diffTemp2 : diffTemp1);
        outS(prepassFollows ? specTemp2 : specTemp1);
        emit("REBLUR_HitDistReconstruction.cs.hlsl" + sigSignal + "|NRD_MODE=RADIANCE" + (is5x5 ? "|MODE_5X5=1" : "|MODE_5X5=0"), 8, 16, kCb);
    }

    for (int i = 0; i < 2; i++) {
        bool afterReconstruction = i & 1;
        beginPass(name("Pre-pass"));
        in(Tr(T_TILES));
        in(U(ResourceType::IN_NORMAL_ROUGHNESS));
        in(U(ResourceType::IN_VIEWZ));
        inD(afterReconstruction ? diffTemp2 : U(inDiff));
        inS(afterReconstruction ? specTemp2 : U(inSpec));
        inShD(U(ResourceType::IN_DIFF_SH1));
        inShS(U(ResourceType::IN_SPEC_SH1));
        outD(diffTemp1);
        outS(specTemp1);
        outS(Tr(T_SPEC_HITDIST_TRACKING));
        outShD(diffShTemp1);
        outShS(specShTemp1);
        emit("REBLUR_PrePass.cs.hlsl" + sig, 16, 16, kCb);
    }

    for (int i = 0; i < 8; i++) {
        bool hasMix = (i >> 2) & 1, hasConfidence = (i >> 1) & 1, afterPrepass = i & 1;
        beginPass(name("Temporal accumulation"));
        in(Tr(T_TILES));
        in(U(ResourceType::IN_NORMAL_ROUGHNESS));
        in(U(ResourceType::IN_VIEWZ));
        in(U(ResourceType::IN_MV));
        in(Pm(P_PREV_VIEWZ));
        in(Pm(P_PREV_NORMAL_ROUGHNESS));
        in(Pm(P_PREV_INTERNAL_DATA));
        in(hasMix ? U(ResourceType::IN_DISOCCLUSION_THRESHOLD_MIX) : dummy);
        inD(hasConfidence ? U(ResourceType::IN_DIFF_CONFIDENCE) : dummy);
        inS(hasConfidence ? U(ResourceType::IN_SPEC_CONFIDENCE) : dummy);
        inD(afterPrepass ? diffTemp1 : U(inDiff));
        inS(afterPrepass ? specTemp1 : U(inSpec));
        inD(Pm(P_DIFF_HISTORY));
        inS(Pm(P_SPEC_HISTORY));
        inD(Pm(P_DIFF_FAST_HISTORY));
        inS(Pm(P_SPEC_FAST_HISTORY));
        inS(Pm(P_SPEC_HITDIST_TRACKING_PING), Pm(P_SPEC_HITDIST_TRACKING_PONG));
        inS(Tr(T_SPEC_HITDIST_TRACKING));
        inShD(afterPrepass ? diffShTemp1 : U(ResourceType::IN_DIFF_SH1));
        inShS(afterPrepass ? specShTemp1 : U(ResourceType::IN_SPEC_SH1));
        inShD(Pm(P_DIFF_SH_HISTORY));
        inShS(Pm(P_SPEC_SH_HISTORY));
        out(Tr(T_DATA1));
        outD(diffTemp2);
        outS(specTemp2);
        outD(Tr(T_DIFF_FAST));
        outS(Tr(T_SPEC_FAST));
        outS(Pm(P_SPEC_HITDIST_TRACKING_PONG), Pm(P_SPEC_HITDIST_TRACKING_PING));
        out(Tr(T_DATA2));
        outShD(diffShTemp2);
        outShS(specShTemp2);
        emit("REBLUR_TemporalAccumulation.cs.hlsl" + sig, 8, 16, kCb);
    }

    beginPass(name("History fix"));
    in(Tr(T_TILES));
    in(U(ResourceType::IN_NORMAL_ROUGHNESS));
    in(Tr(T_DATA1));
    in(U(ResourceType::IN_VIEWZ));
    inD(diffTemp2);
    inS(specTemp2);
    inD(Tr(T_DIFF_FAST));
    inS(Tr(T_SPEC_FAST));
    inS(Pm(P_SPEC_HITDIST_TRACKING_PONG), Pm(P_SPEC_HITDIST_TRACKING_PING));
    inShD(diffShTemp2);
    inShS(specShTemp2);
    outD(diffTemp1);
    outS(specTemp1);
    outD(Pm(P_DIFF_FAST_HISTORY));
    outS(Pm(P_SPEC_FAST_HISTORY));
    outShD(diffShTemp1);
    outShS(specShTemp1);
    emit("REBLUR_HistoryFix.cs.hlsl" + sig, 8, 16, kCb);

    beginPass(name("Blur"));
    in(Tr(T_TILES));
    in(U(ResourceType::IN_NORMAL_ROUGHNESS));
    in(U(ResourceType::IN_VIEWZ));
    in(Tr(T_DATA1));
    inD(diffTemp1);
    inS(specTemp1);
    inShD(diffShTemp1);
    inShS(specShTemp1);
    out(Pm(P_PREV_VIEWZ));
    outD(diffTemp2);
    outS(specTemp2);
    outShD(diffShTemp2);
    outShS(specShTemp2);
    emit("REBLUR_Blur.cs.hlsl" + sig, 8, 16, kCb);

    for (int i = 0; i < 2; i++) {
        bool stabilizationFollows = i & 1;
        beginPass(name("Post-blur"));
        in(Tr(T_TILES));
        in(U(ResourceType::IN_NORMAL_ROUGHNESS));
        in(Tr(T_DATA1));
        in(Pm(P_PREV_VIEWZ));
        inD(diffTemp2);
        inS(specTemp2);
        inShD(diffShTemp2);
        inShS(specShTemp2);
        out(Pm(P_PREV_NORMAL_ROUGHNESS));
        outD(Pm(P_DIFF_HISTORY));
        outS(Pm(P_SPEC_HISTORY));
        if (!stabilizationFollows) {
            out(Pm(P_PREV_INTERNAL_DATA));
            outD(U(outDiff));
            outS(U(outSpec));
            outShD(U(ResourceType::OUT_DIFF_SH1));
            outShS(U(ResourceType::OUT_SPEC_SH1));
        }
        outShD(Pm(P_DIFF_SH_HISTORY));
        outShS(Pm(P_SPEC_SH_HISTORY));
        emit("REBLUR_PostBlur.cs.hlsl" + sig + (stabilizationFollows ? "|TEMPORAL_STABILIZATION=1" : "|TEMPORAL_STABILIZATION=0"), 8, 16, kCb);
    }

    beginPass(name("Temporal stabilization"));
    in(Tr(T_TILES));
    in(U(ResourceType::IN_NORMAL_ROUGHNESS));
    in(Pm(P_PREV_VIEWZ));
    in(Tr(T_DATA1));
    in(Tr(T_DATA2));
    inS(Pm(P_SPEC_HITDIST_TRACKING_PONG), Pm(P_SPEC_HITDIST_TRACKING_PING));
    inD(Pm(P_DIFF_HISTORY));
    inS(Pm(P_SPEC_HISTORY));
    inD(Pm(P_DIFF_STABILIZED_PING), Pm(P_DIFF_STABILIZED_PONG));
    inS(Pm(P_SPEC_STABILIZED_PING), Pm(P_SPEC_STABILIZED_PONG));
    inShD(Pm(P_DIFF_SH_HISTORY));
    inShS(Pm(P_SPEC_SH_HISTORY));
    out(U(ResourceType::IN_MV));  // bound read-write by the reference; only read by this pass
    out(Pm(P_PREV_INTERNAL_DATA));
    outD(U(outDiff));
    outS(U(outSpec));
    outD(Pm(P_DIFF_STABILIZED_PONG), Pm(P_DIFF_STABILIZED_PING));
    outS(Pm(P_SPEC_STABILIZED_PONG), Pm(P_SPEC_STABILIZED_PING));
    outShD(U(ResourceType::OUT_DIFF_SH1));
    outShS(U(ResourceType::OUT_SPEC_SH1));
    emit("REBLUR_TemporalStabilization.cs.hlsl" + sig, 8, 16, kCb);

    beginPass(name("Split screen"));
    in(U(ResourceType::IN_VIEWZ));
    inD(U(inDiff));
    inS(U(inSpec));
    inShD(U(ResourceType::IN_DIFF_SH1));
    inShS(U(ResourceType::IN_SPEC_SH1));
    outD(U(outDiff));
    outS(U(outSpec));
    outShD(U(ResourceType::OUT_DIFF_SH1));
    outShS(U(ResourceType::OUT_SPEC_SH1));
    emit("REBLUR_SplitScreen.cs.hlsl" + (directional ? sigSignal + "|NRD_MODE=RADIANCE" : sig), 8, 16, kCb);

    // REBLUR_ADD_VALIDATION_DISPATCH (Reblur.cpp:65-78): a single-lobe denoiser binds its input in both lobe slots
    beginPass(name("Validation"));
    in(U(ResourceType::IN_NORMAL_ROUGHNESS));
    in(U(ResourceType::IN_VIEWZ));
    in(U(ResourceType::IN_MV));
    in(Tr(T_DATA1));
    in(Tr(T_DATA2));
    in(U(diff ? inDiff : inSpec));
    in(U(spec ? inSpec : inDiff));
    out(U(ResourceType::OUT_VALIDATION));
    emit("REBLUR_Validation.cs.hlsl", 8, 16, kCb, GRID_FROM_RESOURCE, 1);
}

// REBLUR_DIFFUSE_OCCLUSION / REBLUR_SPECULAR_OCCLUSION / REBLUR_DIFFUSE_SPECULAR_OCCLUSION ( NRD_MODE = OCCLUSION; Reblur_DiffuseOcclusion.hpp,
// Reblur_SpecularOcclusion.hpp, Reblur_DiffuseSpecularOcclusion.hpp ): hit distance only — one 16-bit UNORM channel per lobe ( 8-bit fast history ), no pre-pass,
// no data2, no stabilization; the post-blur writes history, internal data and the user's OUT_*_HITDIST. Permanent pool: prev viewZ / normal+roughness / internal
// data, the histories of the lobes, the fast histories of the lobes, the specular tracking ping / pong; transient: data1, per lobe { tmp2, fast history }, tiles.
void Graph::buildReblurOcclusion(DenoiserState& d, bool diff, bool spec) {
    new (&d.settings.reblur) ReblurSettings();
    d.settingsSize = sizeof(ReblurSettings);

    const Format kOcclusion = Format::R16_UNORM, kFast = Format::R8_UNORM;
    uint16_t nPerm = 0, nTran = 0;
    auto perm = [&](Format f) { addPermanent(f); return nPerm++; };
    auto tran = [&](Format f, uint16_t ds = 1) { addTransient(f, ds); return nTran++; };
    const uint16_t P_PREV_VIEWZ = perm(Format::R32_SFLOAT);
    const uint16_t P_PREV_NORMAL_ROUGHNESS = perm(Format::R10_G10_B10_A2_UNORM);
    const uint16_t P_PREV_INTERNAL_DATA = perm(Format::R16_UINT);
    uint16_t P_DIFF_HISTORY = 0, P_SPEC_HISTORY = 0, P_DIFF_FAST_HISTORY = 0, P_SPEC_FAST_HISTORY = 0, P_TRACKING_PING = 0, P_TRACKING_PONG = 0;
    if (diff) P_DIFF_HISTORY = perm(kOcclusion);
    if (spec) P_SPEC_HISTORY = perm(kOcclusion);
    if (diff) P_DIFF_FAST_HISTORY = perm(kFast);
    if (spec) P_SPEC_FAST_HISTORY = perm(kFast);
    if (spec) {
        P_TRACKING_PING = perm(Format::R16_SFLOAT);
        P_TRACKING_PONG = perm(Format::R16_SFLOAT);
    }
    const uint16_t T_DATA1 = tran(diff && spec ? Format::RG8_UNORM : Format::R8_UNORM);
    uint16_t T_DIFF_TMP2 = 0, T_DIFF_FAST = 0, T_SPEC_TMP2 = 0, T_SPEC_FAST = 0;
    if (diff) {
        T_DIFF_TMP2 = tran(kOcclusion);
        T_DIFF_FAST = tran(kFast);
    }
    if (spec) {
        T_SPEC_TMP2 = tran(kOcclusion);
        T_SPEC_FAST = tran(kFast);
    }
    const uint16_t T_TILES = tran(Format::R8_UNORM, 16);

    auto U = [](ResourceType t) { return Slot::user(t); };
    auto Pm = [](uint16_t i) { return Slot::perm(i); };
    auto Tr = [](uint16_t i) { return Slot::tran(i); };
    const Slot inDiff = U(ResourceType::IN_DIFF_HITDIST), inSpec = U(ResourceType::IN_SPEC_HITDIST);
    const Slot diffTemp1 = U(ResourceType::OUT_DIFF_HITDIST), specTemp1 = U(ResourceType::OUT_SPEC_HITDIST);   // the user's outputs double as scratch
    const Slot diffTemp2 = Tr(T_DIFF_TMP2), specTemp2 = Tr(T_SPEC_TMP2);
    const Slot dummy = U(ResourceType::IN_VIEWZ);
    const std::string sigSignal = std::string("|NRD_SIGNAL=") + (diff && spec ? "BOTH" : (diff ? "DIFF" : "SPEC"));
    const std::string sig = sigSignal + "|NRD_MODE=OCCLUSION";
    const std::string prefix = std::string("REBLUR_") + (diff && spec ? "DiffuseSpecular" : (diff ? "Diffuse" : "Specular")) + "Occlusion - ";
    auto name = [&](const char* pass) { return intern(prefix + pass); };
    auto inD = [&](Slot s, Slot swap = Slot()) { if (diff) in(s, swap); };
    auto inS = [&](Slot s, Slot swap = Slot()) { if (spec) in(s, swap); };
    auto outD = [&](Slot s, Slot swap = Slot()) { if (diff) out(s, swap); };
    auto outS = [&](Slot s, Slot swap = Slot()) { if (spec) out(s, swap); };

    beginPass(name("Classify tiles"));
    in(U(ResourceType::IN_VIEWZ));
    out(Tr(T_TILES));
    emit("REBLUR_ClassifyTiles.cs.hlsl", 16, 16, kCb);

    for (int i = 0; i < 2; i++) {
        beginPass(name("Hit distance reconstruction"));
        in(Tr(T_TILES));
        in(U(ResourceType::IN_NORMAL_ROUGHNESS));
        in(U(ResourceType::IN_VIEWZ));
        inD(inDiff);
        inS(inSpec);
        outD(diffTemp1);
        outS(specTemp1);
        emit("REBLUR_HitDistReconstruction.cs.hlsl" + sig + (i ? "|MODE_5X5=1" : "|MODE_5X5=0"), 8, 16, kCb);
    }

    for (int i = 0; i < 8; i++) {
        const bool hasMix = (i >> 2) & 1, hasConfidence = (i >> 1) & 1, afterReconstruction = i & 1;
        beginPass(name("Temporal accumulation"));
        in(Tr(T_TILES));
        in(U(ResourceType::IN_NORMAL_ROUGHNESS));
        in(U(ResourceType::IN_VIEWZ));
        in(U(ResourceType::IN_MV));
        in(Pm(P_PREV_VIEWZ));
        in(Pm(P_PREV_NORMAL_ROUGHNESS));
        in(Pm(P_PREV_INTERNAL_DATA));
        in(hasMix ? U(ResourceType::IN_DISOCCLUSION_THRESHOLD_MIX) : dummy);
        inD(hasConfidence ? U(ResourceType::IN_DIFF_CONFIDENCE) : dummy);
        inS(hasConfidence ? U(ResourceType::IN_SPEC_CONFIDENCE) : dummy);
        inD(afterReconstruction ? diffTemp1 : inDiff);
        inS(afterReconstruction ? specTemp1 : inSpec);
        inD(Pm(P_DIFF_HISTORY));
        inS(Pm(P_SPEC_HISTORY));
        inD(Pm(P_DIFF_FAST_HISTORY));
        inS(Pm(P_SPEC_FAST_HISTORY));
        inS(Pm(P_TRACKING_PING), Pm(P_TRACKING_PONG));
        out(Tr(T_DATA1));
        outD(diffTemp2);
        outS(specTemp2);
        outD(Tr(T_DIFF_FAST));
        outS(Tr(T_SPEC_FAST));
        outS(Pm(P_TRACKING_PONG), Pm(P_TRACKING_PING));
        emit("REBLUR_TemporalAccumulation.cs.hlsl" + sig, 8, 16, kCb);
    }

    beginPass(name("History fix"));
    in(Tr(T_TILES));
    in(U(ResourceType::IN_NORMAL_ROUGHNESS));
    in(Tr(T_DATA1));
    in(U(ResourceType::IN_VIEWZ));
    inD(diffTemp2);
    inS(specTemp2);
    inD(Tr(T_DIFF_FAST));
    inS(Tr(T_SPEC_FAST));
    outD(diffTemp1);
    outS(specTemp1);
    outD(Pm(P_DIFF_FAST_HISTORY));
    outS(Pm(P_SPEC_FAST_HISTORY));
    emit("REBLUR_HistoryFix.cs.hlsl" + sig, 8, 16, kCb);

    beginPass(name("Blur"));
    in(Tr(T_TILES));
    in(U(ResourceType::IN_NORMAL_ROUGHNESS));
    in(U(ResourceType::IN_VIEWZ));
    in(Tr(T_DATA1));
    inD(diffTemp1);
    inS(specTemp1);
    out(Pm(P_PREV_VIEWZ));
    outD(diffTemp2);
    outS(specTemp2);
    emit("REBLUR_Blur.cs.hlsl" + sig, 8, 16, kCb);

    beginPass(name("Post-blur"));
    in(Tr(T_TILES));
    in(U(ResourceType::IN_NORMAL_ROUGHNESS));
    in(Tr(T_DATA1));
    in(Pm(P_PREV_VIEWZ));
    inD(diffTemp2);
    inS(specTemp2);
    out(Pm(P_PREV_NORMAL_ROUGHNESS));
    outD(Pm(P_DIFF_HISTORY));
    outS(Pm(P_SPEC_HISTORY));
    out(Pm(P_PREV_INTERNAL_DATA));
    outD(U(ResourceType::OUT_DIFF_HITDIST));
    outS(U(ResourceType::OUT_SPEC_HITDIST));
    emit("REBLUR_PostBlur.cs.hlsl" + sig + "|TEMPORAL_STABILIZATION=0", 8, 16, kCb);

    beginPass(name("Split screen"));
    in(U(ResourceType::IN_VIEWZ));
    inD(inDiff);
    inS(inSpec);
    outD(U(ResourceType::OUT_DIFF_HITDIST));
    outS(U(ResourceType::OUT_SPEC_HITDIST));
    emit("REBLUR_SplitScreen.cs.hlsl" + sigSignal + "|NRD_MODE=RADIANCE", 8, 16, kCb);

    // REBLUR_ADD_VALIDATION_DISPATCH( Transient::DATA1, ... ): there is no data2, data1 is bound in its slot too
    beginPass(name("Validation"));
    in(U(ResourceType::IN_NORMAL_ROUGHNESS));
    in(U(ResourceType::IN_VIEWZ));
    in(U(ResourceType::IN_MV));
    in(Tr(T_DATA1));
    in(Tr(T_DATA1));
    in(diff ? inDiff : inSpec);
    in(spec ? inSpec : inDiff);
    out(U(ResourceType::OUT_VALIDATION));
    emit("REBLUR_Validation.cs.hlsl", 8, 16, kCb, GRID_FROM_RESOURCE, 1);
}

// Update_ReblurOcclusion ( Reblur.cpp:203-277 )
void Graph::updateReblurOcclusion(const DenoiserState& d) {
    const ReblurSettings& s = d.settings.reblur;
    const bool reconstruct = s.hitDistanceReconstructionMode != HitDistanceReconstructionMode::OFF && s.checkerboardMode == CheckerboardMode::OFF;
    auto push = [&](uint32_t pass) { fillReblurConstants(s, pushDispatch(d, pass)); };
    if (m_common.splitScreen >= 1.0f) {
        push(OCC_SPLIT_SCREEN);
        return;
    }
    push(OCC_CLASSIFY_TILES);
    if (reconstruct) push(OCC_HITDIST_RECONSTRUCTION + (s.hitDistanceReconstructionMode == HitDistanceReconstructionMode::AREA_5X5 ? 1 : 0));
    push(OCC_TEMPORAL_ACCUMULATION + (m_common.isDisocclusionThresholdMixAvailable ? 4 : 0) + (m_common.isHistoryConfidenceAvailable ? 2 : 0) + (reconstruct ? 1 : 0));
    push(OCC_HISTORY_FIX);
    push(OCC_BLUR);
    push(OCC_POST_BLUR);
    if (m_common.splitScreen > 0.0f) push(OCC_SPLIT_SCREEN);
    if (m_common.enableValidation) {
        uint8_t* cb = (uint8_t*)pushDispatch(d, OCC_VALIDATION);
        fillReblurConstants(s, cb);
        uint32_t flags[2] = {hasDiffuse(d.desc.denoiser) ? 1u : 0u, hasSpecular(d.desc.denoiser) ? 1u : 0u};
        memcpy(cb + offsetof(ReblurConstants, _pad), flags, sizeof(flags));
    }
}

void Graph::updateReblur(const DenoiserState& d) {
    const ReblurSettings& s = d.settings.reblur;
    const bool reconstruct = s.hitDistanceReconstructionMode != HitDistanceReconstructionMode::OFF && s.checkerboardMode == CheckerboardMode::OFF;
    const bool skipStabilization = s.maxStabilizedFrameNum == 0;
    const bool diff = hasDiffuse(d.desc.denoiser), spec = hasSpecular(d.desc.denoiser);
    const bool skipPrePass = (s.diffusePrepassBlurRadius == 0.0f || !diff) && (s.specularPrepassBlurRadius == 0.0f || !spec) && s.checkerboardMode == CheckerboardMode::OFF;

    auto push = [&](uint32_t pass) { fillReblurConstants(s, pushDispatch(d, pass)); };

    if (m_common.splitScreen >= 1.0f) {
        push(PASS_SPLIT_SCREEN);
        return;
    }
    push(PASS_CLASSIFY_TILES);
    if (reconstruct)
        push(PASS_HITDIST_RECONSTRUCTION + (s.hitDistanceReconstructionMode == HitDistanceReconstructionMode::AREA_5X5 ? 2 : 0) + (skipPrePass ? 0 : 1));
    if (!skipPrePass) push(PASS_PREPASS + (reconstruct ? 1 : 0));
    push(PASS_TEMPORAL_ACCUMULATION + (m_common.isDisocclusionThresholdMixAvailable ? 4 : 0) + (m_common.isHistoryConfidenceAvailable ? 2 : 0) +
         ((!skipPrePass || reconstruct) ? 1 : 0));
    push(PASS_HISTORY_FIX);
    push(PASS_BLUR);
    push(PASS_POST_BLUR + (skipStabilization ? 0 : 1));
    if (!skipStabilization) push(PASS_TEMPORAL_STABILIZATION);
    if (m_common.splitScreen > 0.0f) push(PASS_SPLIT_SCREEN);
    if (m_common.enableValidation) {
        uint8_t* cb = (uint8_t*)pushDispatch(d, PASS_VALIDATION);
        fillReblurConstants(s, cb);
        // two trailing uints after the shared block: gHasDiffuse, gHasSpecular (the shared block's own tail padding is reused)
        uint32_t flags[2] = {diff ? 1u : 0u, spec ? 1u : 0u};
        memcpy(cb + offsetof(ReblurConstants, _pad), flags, sizeof(flags));
    }
}

void Graph::fillReblurConstants(const ReblurSettings& s, void* dst) {
    if (!dst) return;
    const CommonSettings& c = m_common;
    const FrameState& f = m_frame;
    const float resW = c.resourceSize[0], resH = c.resourceSize[1], resWp = c.resourceSizePrev[0], resHp = c.resourceSizePrev[1];
    const float rectW = c.rectSize[0], rectH = c.rectSize[1], rectWp = c.rectSizePrev[0], rectHp = c.rectSizePrev[1];

    const bool rectChanged = c.rectSize[0] != c.rectSizePrev[0] || c.rectSize[1] != c.rectSizePrev[1];
    const bool reset = c.accumulationMode != AccumulationMode::CONTINUE;
    const float unproject = 1.0f / (0.5f * rectH * f.projectY);
    const float worstScale = std::min(rectW / resW, rectH / resH);
    const float maxBlurRadius = s.maxBlurRadius * worstScale;
    const float thresholdBonus = (1.0f + f.jitterDelta) / rectH;
    const float stabilization = s.maxStabilizedFrameNum / (1.0f + s.maxStabilizedFrameNum);
    const uint32_t maxFrames = std::min(s.maxAccumulatedFrameNum, REBLUR_MAX_HISTORY_FRAME_NUM);

    uint32_t diffCheckerboard = 2, specCheckerboard = 2;
    if (s.checkerboardMode == CheckerboardMode::BLACK) { diffCheckerboard = 0; specCheckerboard = 1; }
    else if (s.checkerboardMode == CheckerboardMode::WHITE) { diffCheckerboard = 1; specCheckerboard = 0; }

    ReblurConstants& k = *(ReblurConstants*)dst;
    k.worldToClip = f.worldToClip;
    k.viewToClip = f.viewToClip;
    k.viewToWorld = f.viewToWorld;
    k.worldToViewPrev = f.worldToViewPrev;
    k.worldToClipPrev = f.worldToClipPrev;
    k.worldPrevToWorld = f.worldPrevToWorld;
    memcpy(k.rotatorPre, f.rotatorPre, 16);
    memcpy(k.rotator, f.rotator, 16);
    memcpy(k.rotatorPost, f.rotatorPost, 16);
    memcpy(k.frustum, f.frustum, 16);
    memcpy(k.frustumPrev, f.frustumPrev, 16);
    for (int i = 0; i < 3; i++) {
        k.cameraDelta[i] = f.cameraDelta[i];
        k.viewVectorWorld[i] = f.viewDirection[i];
        k.viewVectorWorldPrev[i] = f.viewDirectionPrev[i];
        k.mvScale[i] = c.motionVectorScale[i];
    }
    // the reference negates the whole SSE register of the matrix column ( InstanceImpl.cpp:434-435 ): .w = -( 0 ) = -0.0f, never read by a shader
    k.viewVectorWorld[3] = -f.viewToWorld.m[11];
    k.viewVectorWorldPrev[3] = -f.viewToWorldPrev.m[11];
    k.hitDistSettings[0] = s.hitDistanceParameters.A;
    k.hitDistSettings[1] = s.hitDistanceParameters.B;
    k.hitDistSettings[2] = s.hitDistanceParameters.C;
    k.mvScale[3] = c.isMotionVectorInWorldSpace ? 1.0f : 0.0f;
    k.convergenceSettings[0] = s.convergenceSettings.s;
    k.convergenceSettings[1] = s.convergenceSettings.b;
    k.convergenceSettings[2] = s.convergenceSettings.p;
    k.antilagSettings[0] = s.antilagSettings.luminanceSigmaScale;
    k.antilagSettings[1] = s.antilagSettings.luminanceSensitivity;
    k.resourceSize[0] = resW; k.resourceSize[1] = resH;
    k.resourceSizeInv[0] = 1.0f / resW; k.resourceSizeInv[1] = 1.0f / resH;
    k.resourceSizeInvPrev[0] = 1.0f / resWp; k.resourceSizeInvPrev[1] = 1.0f / resHp;
    k.rectSize[0] = rectW; k.rectSize[1] = rectH;
    k.rectSizeInv[0] = 1.0f / rectW; k.rectSizeInv[1] = 1.0f / rectH;
    k.rectSizePrev[0] = rectWp; k.rectSizePrev[1] = rectHp;
    k.resolutionScale[0] = rectW / resW; k.resolutionScale[1] = rectH / resH;
    k.resolutionScalePrev[0] = rectWp / resWp; k.resolutionScalePrev[1] = rectHp / resHp;
    k.rectOffset[0] = float(c.rectOrigin[0]) / resW; k.rectOffset[1] = float(c.rectOrigin[1]) / resH;
    k.jitter[0] = c.cameraJitter[0]; k.jitter[1] = c.cameraJitter[1];
    k.printfAt[0] = c.printfAt[0]; k.printfAt[1] = c.printfAt[1];
    k.rectOrigin[0] = c.rectOrigin[0]; k.rectOrigin[1] = c.rectOrigin[1];
    k.rectSizeMinusOne[0] = c.rectSize[0] - 1; k.rectSizeMinusOne[1] = c.rectSize[1] - 1;
    k.disocclusionThreshold = c.disocclusionThreshold + thresholdBonus;
    k.disocclusionThresholdAlternate = c.disocclusionThresholdAlternate + thresholdBonus;
    k.cameraAttachedReflectionMaterialID = c.cameraAttachedReflectionMaterialID;
    k.strandMaterialID = c.strandMaterialID;
    k.strandThickness = c.strandThickness;
    k.stabilizationStrength = reset ? 0.0f : stabilization;
    k.debug = c.debug;
    k.orthoMode = f.orthoMode;
    k.unproject = unproject;
    k.denoisingRange = c.denoisingRange;
    k.planeDistSensitivity = s.planeDistanceSensitivity;
    k.framerateScale = f.frameRateScale;
    k.maxBlurRadius = std::max(maxBlurRadius, s.minBlurRadius);
    k.minBlurRadius = s.minBlurRadius;
    k.diffPrepassBlurRadius = s.diffusePrepassBlurRadius * worstScale;
    k.specPrepassBlurRadius = s.specularPrepassBlurRadius * worstScale;
    k.maxAccumulatedFrameNum = reset ? 0.0f : float(maxFrames);
    k.maxFastAccumulatedFrameNum = reset ? 0.0f : float(s.maxFastAccumulatedFrameNum);
    k.antiFirefly = s.enableAntiFirefly ? 1.0f : 0.0f;
    k.lobeAngleFraction = s.lobeAngleFraction * s.lobeAngleFraction;  // squared on purpose (Reblur.cpp:367)
    k.roughnessFraction = s.roughnessFraction;
    k.historyFixFrameNum = (float)s.historyFixFrameNum;
    k.historyFixBasePixelStride = (float)s.historyFixBasePixelStride;
    k.historyFixAlternatePixelStride = (float)s.historyFixAlternatePixelStride;
    k.historyFixAlternatePixelStrideMaterialID = c.historyFixAlternatePixelStrideMaterialID;
    {
        float t = std::max(maxBlurRadius, s.minBlurRadius) / 2.0f;
        t = std::min(std::max(t, 0.0f), 1.0f);
        k.fastHistoryClampingSigmaScale = 3.0f + (s.fastHistoryClampingSigmaScale - 3.0f) * t;
    }
    k.minRectDimMulUnproject = (float)std::min(c.rectSize[0], c.rectSize[1]) * unproject;
    k.usePrepassNotOnlyForSpecularMotionEstimation = s.usePrepassOnlyForSpecularMotionEstimation ? 0.0f : 1.0f;
    k.splitScreen = c.splitScreen;
    k.splitScreenPrev = f.splitScreenPrev;
    k.checkerboardResolveAccumSpeed = f.checkerboardResolveAccumSpeed;
    k.viewZScale = c.viewZScale;
    k.fireflySuppressorMinRelativeScale = s.fireflySuppressorMinRelativeScale;
    k.minHitDistanceWeight = s.minHitDistanceWeight;
    k.diffMinMaterial = s.minMaterialForDiffuse;
    k.specMinMaterial = s.minMaterialForSpecular;
    k.responsiveAccumulationInvRoughnessThreshold = 1.0f / std::max(s.responsiveAccumulationSettings.roughnessThreshold, 1e-3f);
    k.responsiveAccumulationMinAccumulatedFrameNum = s.responsiveAccumulationSettings.minAccumulatedFrameNum;
    k.hasHistoryConfidence = c.isHistoryConfidenceAvailable;
    k.hasDisocclusionThresholdMix = c.isDisocclusionThresholdMixAvailable;
    k.diffCheckerboard = diffCheckerboard;
    k.specCheckerboard = specCheckerboard;
    k.frameIndex = c.frameIndex;
    k.isRectChanged = rectChanged ? 1 : 0;
    k.resetHistory = reset ? 1 : 0;
    k.returnHistoryLengthInsteadOfOcclusion = s.returnHistoryLengthInsteadOfOcclusion ? 1 : 0;
}

}  // namespace nrdb
