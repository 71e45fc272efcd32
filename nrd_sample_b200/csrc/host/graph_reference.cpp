// REFERENCE denoiser (plain temporal accumulation of IN_SIGNAL into an RGBA32F history, used to produce ground truth).
// Pool layout and bindings: External/NRD/Source/Denoisers/Reference.hpp:13-83. The accumulated frame count restarts when the
// camera moves, the accumulation mode is not CONTINUE or the rect changes (Reference.hpp:62-68).
#include <algorithm>
#include <cstring>

#include "pass_graph.h"

namespace nrdb {

namespace {
enum PassIndex : uint32_t { PASS_ACCUMULATE, PASS_COPY };
constexpr uint32_t REFERENCE_MAX_HISTORY_FRAME_NUM = 4095;  // NRDSettings.h:480
}  // namespace

void Graph::buildReference(DenoiserState& d) {
    new (&d.settings.reference) ReferenceSettings();
    d.settingsSize = sizeof(ReferenceSettings);

    addPermanent(Format::RGBA32_SFLOAT);  // history

    beginPass("Reference - Temporal accumulation");
    in(Slot::user(ResourceType::IN_SIGNAL));
    out(Slot::perm(0));
    emit("REFERENCE_TemporalAccumulation.cs.hlsl", 16, 16, sizeof(ReferenceAccumulateConstants));

    beginPass("Reference - Copy");
    in(Slot::perm(0));
    out(Slot::user(ResourceType::OUT_SIGNAL));
    emit("REFERENCE_Copy.cs.hlsl", 16, 16, sizeof(ReferenceCopyConstants));
}

void Graph::updateReference(const DenoiserState& d) {
    const ReferenceSettings& s = d.settings.reference;
    const CommonSettings& c = m_common;
    bool cameraMoved = false;  // float4x4::operator!= (MathLib Guts/f32.h:925): any component differs
    for (int i = 0; i < 16; i++) cameraMoved |= m_frame.worldToClip.m[i] != m_frame.worldToClipPrev.m[i];
    if (cameraMoved || c.accumulationMode != AccumulationMode::CONTINUE || c.rectSize[0] != c.rectSizePrev[0] || c.rectSize[1] != c.rectSizePrev[1])
        m_accumulatedFrameNum = 0;
    else
        m_accumulatedFrameNum = std::min(m_accumulatedFrameNum + 1, std::min(s.maxAccumulatedFrameNum, REFERENCE_MAX_HISTORY_FRAME_NUM));

    if (auto* k = (ReferenceAccumulateConstants*)pushDispatch(d, PASS_ACCUMULATE)) {
        k->accumSpeed = 1.0f / (1.0f + (float)m_accumulatedFrameNum);
        k->debug = c.debug;
    }
    if (auto* k = (ReferenceCopyConstants*)pushDispatch(d, PASS_COPY)) {
        k->rectSizeInv[0] = 1.0f / float(c.rectSize[0]);
        k->rectSizeInv[1] = 1.0f / float(c.rectSize[1]);
        k->splitScreen = c.splitScreen;
    }
}

}  // namespace nrdb
