// SIGMA_SHADOW / SIGMA_SHADOW_TRANSLUCENCY pass graph and per-frame constants.
// Pool layout and bindings: External/NRD/Source/Denoisers/Sigma_Shadow.hpp:13-165, Sigma_ShadowTranslucency.hpp:13-168
// (RGBA8 shadow+translucency textures, IN_TRANSLUCENCY bound to ClassifyTiles / Blur / SplitScreen, TRANSLUCENCY=1 permutations).
// Per-frame pass selection: External/NRD/Source/Sigma.cpp:25-85. Constants: Sigma.cpp:87-140.
#include <algorithm>
#include <cmath>
#include <string>

#include "pass_graph.h"

namespace nrdb {

namespace {
enum P : uint16_t { P_HISTORY_LENGTH };
enum T : uint16_t { T_DATA_1, T_DATA_2, T_TEMP_1, T_TEMP_2, T_HISTORY, T_HISTORY_LENGTH, T_TILES, T_SMOOTHED_TILES };
enum PassIndex : uint32_t {
    PASS_CLASSIFY_TILES,
    PASS_SMOOTH_TILES,
    PASS_COPY,
    PASS_BLUR,
    PASS_POST_BLUR,  // 2 permutations: bit0 = temporal stabilization follows
    PASS_TEMPORAL_STABILIZATION = PASS_POST_BLUR + 2,
    PASS_SPLIT_SCREEN,
};
const uint32_t kCb = sizeof(SigmaConstants);
}  // namespace

void Graph::buildSigmaShadow(DenoiserState& d, bool translucency) {
    const Format shadowFormat = translucency ? Format::RGBA8_UNORM : Format::R8_UNORM;
    const std::string tr = translucency ? "|TRANSLUCENCY=1" : "|TRANSLUCENCY=0";
    const std::string name = translucency ? "SIGMA_ShadowTranslucency" : "SIGMA_Shadow";
    new (&d.settings.sigma) SigmaSettings();
    d.settingsSize = sizeof(SigmaSettings);

    addPermanent(Format::R32_UINT);  // asuint(viewZ) with the history length in the low 3 bits

    addTransient(Format::R16_SFLOAT);       // penumbra ping
    addTransient(Format::R16_SFLOAT);       // penumbra pong
    addTransient(shadowFormat);             // shadow ( + translucency ) ping
    addTransient(shadowFormat);             // shadow ( + translucency ) pong
    addTransient(shadowFormat);             // history copy
    addTransient(Format::R32_UINT);         // history-length copy
    addTransient(Format::RGBA8_UNORM, 16);  // tiles
    addTransient(Format::RG8_UNORM, 16);    // smoothed tiles

    auto U = [](ResourceType t) { return Slot::user(t); };
    auto Pm = [](uint16_t i) { return Slot::perm(i); };
    auto Tr = [](uint16_t i) { return Slot::tran(i); };

    beginPass(intern(name + " - Classify tiles"));
    in(U(ResourceType::IN_VIEWZ));
    in(U(ResourceType::IN_PENUMBRA));
    if (translucency) in(U(ResourceType::IN_TRANSLUCENCY));
    out(Tr(T_TILES));
    emit(("SIGMA_ClassifyTiles.cs.hlsl" + tr).c_str(), 16, 16, kCb);

    beginPass(intern(name + " - Smooth tiles"));
    in(Tr(T_TILES));
    out(Tr(T_SMOOTHED_TILES));
    emit("SIGMA_SmoothTiles.cs.hlsl", 16, 16, kCb, 16, 1);

    beginPass(intern(name + " - Copy"));
    in(Tr(T_SMOOTHED_TILES));
    in(U(ResourceType::OUT_SHADOW_TRANSLUCENCY));
    in(Pm(P_HISTORY_LENGTH));
    out(Tr(T_HISTORY));
    out(Tr(T_HISTORY_LENGTH));
    emit("SIGMA_Copy.cs.hlsl", 8, 16, kCb, GRID_FROM_PREV_RECT, 1);

    beginPass(intern(name + " - Blur"));
    in(U(ResourceType::IN_VIEWZ));
    in(U(ResourceType::IN_NORMAL_ROUGHNESS));
    in(U(ResourceType::IN_PENUMBRA));
    in(Tr(T_SMOOTHED_TILES));
    if (translucency) in(U(ResourceType::IN_TRANSLUCENCY));
    out(Tr(T_DATA_1));
    out(Tr(T_TEMP_1));
    emit(("SIGMA_Blur.cs.hlsl" + tr + "|FIRST_PASS=1").c_str(), 8, 16, kCb);

    for (int i = 0; i < 2; i++) {
        bool stabilizationFollows = i & 1;
        beginPass(intern(name + " - Post-blur"));
        in(U(ResourceType::IN_VIEWZ));
        in(U(ResourceType::IN_NORMAL_ROUGHNESS));
        in(Tr(T_DATA_1));
        in(Tr(T_SMOOTHED_TILES));
        in(Tr(T_TEMP_1));
        out(Tr(T_DATA_2));
        out(stabilizationFollows ? Tr(T_TEMP_2) : U(ResourceType::OUT_SHADOW_TRANSLUCENCY));
        emit(("SIGMA_Blur.cs.hlsl" + tr + "|FIRST_PASS=0").c_str(), 8, 16, kCb);
    }

    beginPass(intern(name + " - Temporal stabilization"));
    in(U(ResourceType::IN_VIEWZ));
    in(U(ResourceType::IN_MV));
    in(Tr(T_DATA_2));
    in(Tr(T_TEMP_2));
    in(Tr(T_HISTORY));
    in(Tr(T_HISTORY_LENGTH));
    in(Tr(T_SMOOTHED_TILES));
    out(U(ResourceType::OUT_SHADOW_TRANSLUCENCY));
    out(Pm(P_HISTORY_LENGTH));
    emit(("SIGMA_TemporalStabilization.cs.hlsl" + tr).c_str(), 8, 16, kCb);

    beginPass(intern(name + " - Split screen"));
    in(U(ResourceType::IN_VIEWZ));
    in(U(ResourceType::IN_PENUMBRA));
    if (translucency) in(U(ResourceType::IN_TRANSLUCENCY));
    out(U(ResourceType::OUT_SHADOW_TRANSLUCENCY));
    emit(("SIGMA_SplitScreen.cs.hlsl" + tr).c_str(), 8, 16, kCb);
}

void Graph::updateSigma(const DenoiserState& d) {
    const SigmaSettings& s = d.settings.sigma;
    auto push = [&](uint32_t pass) { fillSigmaConstants(s, pushDispatch(d, pass)); };
    if (m_common.splitScreen >= 1.0f) {
        push(PASS_SPLIT_SCREEN);
        return;
    }
    push(PASS_CLASSIFY_TILES);
    push(PASS_SMOOTH_TILES);
    if (s.maxStabilizedFrameNum) push(PASS_COPY);
    push(PASS_BLUR);
    push(PASS_POST_BLUR + (s.maxStabilizedFrameNum ? 1 : 0));
    if (s.maxStabilizedFrameNum) push(PASS_TEMPORAL_STABILIZATION);
    if (m_common.splitScreen > 0.0f) push(PASS_SPLIT_SCREEN);
}

void Graph::fillSigmaConstants(const SigmaSettings& s, void* dst) {
    if (!dst) return;
    const CommonSettings& c = m_common;
    const FrameState& f = m_frame;
    const float resW = c.resourceSize[0], resH = c.resourceSize[1], resWp = c.resourceSizePrev[0], resHp = c.resourceSizePrev[1];
    const float rectW = c.rectSize[0], rectH = c.rectSize[1];
    const float unproject = 1.0f / (0.5f * rectH * f.projectY);
    const uint32_t frameNum = std::min(s.maxStabilizedFrameNum, SIGMA_MAX_HISTORY_FRAME_NUM);
    const float stabilization = frameNum / (1.0f + frameNum);

    SigmaConstants& k = *(SigmaConstants*)dst;
    k.worldToView = f.worldToView;
    k.viewToClip = f.viewToClip;
    k.worldToClipPrev = f.worldToClipPrev;
    k.worldToViewPrev = f.worldToViewPrev;
    memcpy(k.rotator, f.rotator, 16);
    memcpy(k.rotatorPost, f.rotatorPost, 16);
    memcpy(k.frustum, f.frustum, 16);
    memcpy(k.frustumPrev, f.frustumPrev, 16);
    k.viewVectorWorld[3] = -f.viewToWorld.m[11];   // the reference negates the whole SSE register ( InstanceImpl.cpp:434 ): -0.0f
    for (int i = 0; i < 3; i++) {
        k.viewVectorWorld[i] = f.viewDirection[i];
        k.cameraDelta[i] = f.cameraDelta[i];
        k.mvScale[i] = c.motionVectorScale[i];
        // light direction rotated into view space
        const float* L = s.lightDirection;
        float v = L[0] * f.worldToView.m[0 * 4 + i];
        v = L[1] * f.worldToView.m[1 * 4 + i] + v;
        v = L[2] * f.worldToView.m[2 * 4 + i] + v;
        k.lightDirectionView[i] = v;
    }
    k.mvScale[3] = c.isMotionVectorInWorldSpace ? 1.0f : 0.0f;
    k.resourceSizeInv[0] = 1.0f / resW; k.resourceSizeInv[1] = 1.0f / resH;
    k.resourceSizeInvPrev[0] = 1.0f / resWp; k.resourceSizeInvPrev[1] = 1.0f / resHp;
    k.rectSize[0] = rectW; k.rectSize[1] = rectH;
    k.rectSizeInv[0] = 1.0f / rectW; k.rectSizeInv[1] = 1.0f / rectH;
    k.rectSizePrev[0] = c.rectSizePrev[0]; k.rectSizePrev[1] = c.rectSizePrev[1];
    k.resolutionScale[0] = rectW / resW; k.resolutionScale[1] = rectH / resH;
    k.rectOffset[0] = float(c.rectOrigin[0]) / resW; k.rectOffset[1] = float(c.rectOrigin[1]) / resH;
    k.printfAt[0] = c.printfAt[0]; k.printfAt[1] = c.printfAt[1];
    k.rectOrigin[0] = c.rectOrigin[0]; k.rectOrigin[1] = c.rectOrigin[1];
    k.rectSizeMinusOne[0] = c.rectSize[0] - 1; k.rectSizeMinusOne[1] = c.rectSize[1] - 1;
    k.tilesSizeMinusOne[0] = (c.rectSize[0] + 15) / 16 - 1; k.tilesSizeMinusOne[1] = (c.rectSize[1] + 15) / 16 - 1;
    k.orthoMode = f.orthoMode;
    k.unproject = unproject;
    k.denoisingRange = c.denoisingRange;
    k.planeDistSensitivity = s.planeDistanceSensitivity;
    k.stabilizationStrength = c.accumulationMode == AccumulationMode::CONTINUE ? stabilization : 0.0f;
    k.debug = c.debug;
    k.splitScreen = c.splitScreen;
    k.viewZScale = c.viewZScale;
    k.minRectDimMulUnproject = (float)std::min(c.rectSize[0], c.rectSize[1]) * unproject;
    k.frameIndex = c.frameIndex;
    k.isRectChanged = (c.rectSize[0] != c.rectSizePrev[0] || c.rectSize[1] != c.rectSizePrev[1]) ? 1 : 0;
}

}  // namespace nrdb
