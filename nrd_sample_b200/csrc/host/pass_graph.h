// Host-side pass graph of the B200 NRD replacement: what runs when, on which texture, with which
// constants. No GPU work happens here — this is the part of the reference that lives in
// External/NRD/Source/InstanceImpl.{h,cpp} and it must reproduce that library's DispatchDesc
// streams (names, order, bindings, ping-pong parity, grid sizes, constant bytes) for the denoisers
// on the hot path. tests/test_dispatch_stream.py diffs the two libraries field by field.
#pragma once

#include <cstring>
#include <deque>
#include <string>
#include <vector>

#include "../../../include/nrd_b200.h"
#include "constants.h"

namespace nrdb {

using namespace nrd;

// A texture named inside one denoiser's graph, before pools are merged into the instance.
struct Slot {
    enum Kind : uint8_t { USER, PERMANENT, TRANSIENT, NONE } kind = NONE;
    uint16_t index = 0;  // ResourceType value for USER, local pool index otherwise
    static Slot user(ResourceType t) { return {USER, (uint16_t)t}; }
    static Slot perm(uint16_t i) { return {PERMANENT, i}; }
    static Slot tran(uint16_t i) { return {TRANSIENT, i}; }
};

// Sentinels for the grid computation (InstanceImpl.h:154-155).
constexpr uint16_t GRID_FROM_PREV_RECT = 0xFFFF;
constexpr uint16_t GRID_FROM_RESOURCE = 0xFFFE;

struct PassRecord {
    const char* name;          // static string, shared by permutations of one pass
    size_t firstResource;      // into Graph::resources
    uint32_t resourcesNum;
    uint32_t constantsSize;
    Identifier identifier;
    uint16_t pipelineIndex;
    uint16_t downsample;       // or one of the GRID_FROM_* sentinels
    uint16_t maxRepeat;
    uint8_t groupW, groupH;
};

struct SwapEntry {
    size_t resourceIndex;      // binding whose pool index flips every frame
    uint16_t partner;
};

struct ClearTarget {
    Identifier identifier;
    ResourceDesc resource;
    uint16_t downsample;
    bool isInteger;
};

union AnySettings {
    ReblurSettings reblur;
    RelaxSettings relax;
    SigmaSettings sigma;
    ReferenceSettings reference;
    AnySettings() {}
};

struct DenoiserState {
    DenoiserDesc desc;
    AnySettings settings;
    size_t settingsSize = 0;
    size_t firstPass = 0;
    size_t firstSwap = 0, swapNum = 0;
};

// Per-frame camera state derived in SetCommonSettings (InstanceImpl.cpp:330-451).
struct FrameState {
    Mat4 viewToClip, viewToClipPrev, worldToView, worldToViewPrev, viewToWorld, viewToWorldPrev;
    Mat4 worldToClip, worldToClipPrev, worldPrevToWorld;
    float rotatorPre[4], rotator[4], rotatorPost[4];
    float frustum[4], frustumPrev[4];
    float cameraDelta[3], viewDirection[3], viewDirectionPrev[3];
    float splitScreenPrev = 0.0f, orthoMode = 0.0f, checkerboardResolveAccumSpeed = 0.0f, jitterDelta = 0.0f;
    float timeDelta = 0.0f, frameRateScale = 0.0f, projectY = 0.0f;
};

class Graph {
public:
    Graph();
    Result create(const InstanceCreationDesc& desc);
    Result setCommonSettings(const CommonSettings& cs);
    Result setDenoiserSettings(Identifier id, const void* settings);
    Result getComputeDispatches(const Identifier* ids, uint32_t idsNum, const DispatchDesc*& out, uint32_t& outNum);
    const InstanceDesc& desc() const { return m_desc; }

    // ---- builder interface used by graph_*.cpp --------------------------------------------
    void addPermanent(Format f, uint16_t downsample = 1) { m_permanentPool.push_back({f, downsample}); }
    void addTransient(Format f, uint16_t downsample = 1);
    void beginPass(const char* name) {
        m_passName = name;
        m_passFirstResource = m_resources.size();
    }
    void in(Slot s, Slot swapWith = Slot()) { bind(DescriptorType::TEXTURE, s, swapWith); }
    void out(Slot s, Slot swapWith = Slot()) { bind(DescriptorType::STORAGE_TEXTURE, s, swapWith); }
    // shaderId: "File.cs.hlsl|A=1|B=2" exactly as InstanceImpl.h:59-67 would print it
    void emit(const std::string& shaderId, uint8_t groupW, uint8_t groupH, uint32_t constantsSize, uint16_t downsample = 1,
              uint16_t maxRepeat = 1);

    // ---- per-frame interface used by graph_*.cpp ------------------------------------------
    void* pushDispatch(const DenoiserState& d, uint32_t localPassIndex);
    const CommonSettings& common() const { return m_common; }
    const FrameState& frame() const { return m_frame; }

private:
    void bind(DescriptorType type, Slot s, Slot swapWith);
    void finalize();
    void flipPingPong(const DenoiserState& d);

    // graph_reblur.cpp / graph_sigma.cpp / graph_relax.cpp
    void buildReblur(DenoiserState& d, bool diff, bool spec, bool sh = false, bool directional = false);
    void updateReblur(const DenoiserState& d);
    void buildReblurOcclusion(DenoiserState& d, bool diff, bool spec);
    void updateReblurOcclusion(const DenoiserState& d);
    void fillReblurConstants(const ReblurSettings& s, void* dst);
    void buildSigmaShadow(DenoiserState& d, bool translucency);
    void updateSigma(const DenoiserState& d);
    void fillSigmaConstants(const SigmaSettings& s, void* dst);
    void buildRelax(DenoiserState& d, bool diff, bool spec, bool sh);
    void updateRelax(const DenoiserState& d);
    void* fillRelaxConstants(const RelaxSettings& s, void* dst);
    void buildReference(DenoiserState& d);
    void updateReference(const DenoiserState& d);

    std::vector<DenoiserState> m_denoisers;
    std::vector<TextureDesc> m_permanentPool, m_transientPool;
    std::vector<ResourceDesc> m_resources;
    std::vector<ClearTarget> m_clears;
    std::vector<SwapEntry> m_swaps;
    std::vector<ResourceRangeDesc> m_ranges;
    std::vector<size_t> m_pipelineFirstRange;
    std::vector<PipelineDesc> m_pipelines;
    std::vector<PassRecord> m_passes;
    std::vector<DispatchDesc> m_active;
    std::vector<uint16_t> m_transientRemap;  // local transient index -> instance pool index, per denoiser being built
    std::vector<uint8_t> m_constantArena;
    uint8_t* m_constantData = nullptr;
    size_t m_constantOffset = 0;
    size_t m_clearPass[2] = {};
    std::deque<std::string> m_names;   // pass names built at run time ( interned: finalize() groups permutations by name pointer )
    const char* intern(const std::string& s) {
        for (const std::string& n : m_names) if (n == s) return n.c_str();
        m_names.push_back(s);
        return m_names.back().c_str();
    }
    const char* m_passName = nullptr;
    size_t m_passFirstResource = 0;
    uint16_t m_permanentBase = 0, m_transientBase = 0;
    InstanceDesc m_desc = {};
    CommonSettings m_common = {};
    FrameState m_frame = {};
    uint32_t m_accumulatedFrameNum = 0;  // REFERENCE denoiser (InstanceImpl::m_AccumulatedFrameNum)
    bool m_firstUse = true;
    double m_lastTimeMs = -1.0;
    float m_smoothedDeltaMs = 16.6f;
};

const LibraryDesc& libraryDesc();

}  // namespace nrdb
