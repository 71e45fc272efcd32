// Pass-graph bookkeeping: pools, bindings, ping-pong, clear injection, per-frame camera state, and
// the nine exported entry points. Behaviour follows External/NRD/Source/InstanceImpl.cpp:88-823 and
// Wrapper.cpp:182-239; the data structures are this library's own.
#include "pass_graph.h"

#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <new>

#include "mat.h"

namespace nrdb {

static constexpr size_t kConstantArenaSize = 128 * 1024;  // InstanceImpl.h:150

static const Sampler kSamplers[2] = {Sampler::NEAREST_CLAMP, Sampler::LINEAR_CLAMP};

static bool isIntegerFormat(Format f) {
    switch (f) {
        case Format::R8_UINT: case Format::RG8_UINT: case Format::RGBA8_UINT: case Format::R16_UINT: case Format::RG16_UINT:
        case Format::RGBA16_UINT: case Format::R32_UINT: case Format::RG32_UINT: case Format::RGB32_UINT: case Format::RGBA32_UINT:
        case Format::R10_G10_B10_A2_UINT:
            return true;  // same set as g_IsIntegerFormat (InstanceImpl.cpp:18-63): *_SINT formats count as float there
        default:
            return false;
    }
}

static uint16_t divUp(uint32_t x, uint16_t y) { return (uint16_t)((x + y - 1) / y); }

static bool contains(const Identifier* ids, uint32_t n, Identifier id) {
    for (uint32_t i = 0; i < n; i++)
        if (ids[i] == id) return true;
    return false;
}

// The denoisers this build implements end to end (graph + CUDA kernels). Everything else is UNSUPPORTED,
// which is the reference's own answer for a denoiser missing from LibraryDesc (InstanceImpl.cpp:95-102).
static const Denoiser kSupported[] = {
    Denoiser::REBLUR_DIFFUSE,
    Denoiser::REBLUR_SPECULAR,
    Denoiser::REBLUR_DIFFUSE_SPECULAR,
    Denoiser::REBLUR_DIFFUSE_SH,
    Denoiser::REBLUR_SPECULAR_SH,
    Denoiser::REBLUR_DIFFUSE_SPECULAR_SH,
    Denoiser::REBLUR_DIFFUSE_OCCLUSION,
    Denoiser::REBLUR_SPECULAR_OCCLUSION,
    Denoiser::REBLUR_DIFFUSE_SPECULAR_OCCLUSION,
    Denoiser::REBLUR_DIFFUSE_DIRECTIONAL_OCCLUSION,
    Denoiser::RELAX_DIFFUSE,
    Denoiser::RELAX_DIFFUSE_SH,
    Denoiser::RELAX_SPECULAR,
    Denoiser::RELAX_SPECULAR_SH,
    Denoiser::RELAX_DIFFUSE_SPECULAR,
    Denoiser::RELAX_DIFFUSE_SPECULAR_SH,
    Denoiser::SIGMA_SHADOW,
    Denoiser::SIGMA_SHADOW_TRANSLUCENCY,
    Denoiser::REFERENCE,
};

const LibraryDesc& libraryDesc() {
    static const LibraryDesc d = {
        {0, 20, 2, 3},  // SPIRV sampler/texture/cb/uav offsets (NRD/CMakeLists.txt:90-94) — informational only
        kSupported,
        (uint32_t)(sizeof(kSupported) / sizeof(kSupported[0])),
        NRD_VERSION_MAJOR,
        NRD_VERSION_MINOR,
        NRD_VERSION_BUILD,
        NormalEncoding::R10_G10_B10_A2_UNORM,
        RoughnessEncoding::LINEAR,
    };
    return d;
}

Graph::Graph() {
    m_constantArena.resize(kConstantArenaSize + 16);
    uintptr_t p = (uintptr_t)m_constantArena.data();
    m_constantData = (uint8_t*)((p + 15) & ~(uintptr_t)15);
    memset(m_constantData, 0, kConstantArenaSize);
    m_frame.viewToClip = m_frame.viewToClipPrev = m_frame.worldToView = m_frame.worldToViewPrev = mat::identity();
}

// ------------------------------------------------------------------------------------------------
// Builder
// ------------------------------------------------------------------------------------------------
void Graph::addTransient(Format f, uint16_t downsample) {
    // Transient textures of earlier denoisers in the same instance are reused when format and size match
    // and this denoiser does not hold them already (InstanceImpl.cpp:745-770).
    for (uint16_t i = 0; i < m_transientBase; i++) {
        const TextureDesc& t = m_transientPool[i];
        if (t.format != f || t.downsampleFactor != downsample) continue;
        bool taken = false;
        for (uint16_t r : m_transientRemap) taken |= (r == i);
        if (!taken) {
            m_transientRemap.push_back(i);
            return;
        }
    }
    m_transientRemap.push_back((uint16_t)m_transientPool.size());
    m_transientPool.push_back({f, downsample});
}

void Graph::bind(DescriptorType type, Slot s, Slot swapWith) {
    ResourceDesc r = {type, ResourceType::MAX_NUM, 0};
    auto resolve = [&](Slot x) -> uint16_t {
        return x.kind == Slot::TRANSIENT ? m_transientRemap[x.index] : (uint16_t)(m_permanentBase + x.index);
    };
    if (s.kind == Slot::USER) {
        r.type = (ResourceType)s.index;
    } else {
        r.type = s.kind == Slot::TRANSIENT ? ResourceType::TRANSIENT_POOL : ResourceType::PERMANENT_POOL;
        r.indexInPool = resolve(s);
        if (swapWith.kind != Slot::NONE) m_swaps.push_back({m_resources.size(), resolve(swapWith)});
    }
    m_resources.push_back(r);
}

void Graph::emit(const std::string& shaderId, uint8_t groupW, uint8_t groupH, uint32_t constantsSize, uint16_t downsample,
                 uint16_t maxRepeat) {
    // One pipeline per emitted pass: with no embedded bytecode the reference never dedups either
    // (the #if chain at InstanceImpl.cpp:581-590 is empty), so pipeline indices line up.
    PipelineDesc p = {};
    snprintf(p.shaderIdentifier, sizeof(p.shaderIdentifier), "%s", shaderId.c_str());
    p.hasConstantData = constantsSize != 0;
    m_pipelineFirstRange.push_back(m_ranges.size());
    for (int pass = 0; pass < 2; pass++) {
        ResourceRangeDesc range = {pass == 0 ? DescriptorType::TEXTURE : DescriptorType::STORAGE_TEXTURE, 0};
        for (size_t i = m_passFirstResource; i < m_resources.size(); i++) range.descriptorsNum += m_resources[i].descriptorType == range.descriptorType;
        if (range.descriptorsNum) {
            m_ranges.push_back(range);
            p.resourceRangesNum++;
        }
    }
    m_pipelines.push_back(p);

    PassRecord rec = {};
    rec.name = m_passName;
    rec.firstResource = m_passFirstResource;
    rec.resourcesNum = (uint32_t)(m_resources.size() - m_passFirstResource);
    rec.constantsSize = constantsSize;
    rec.pipelineIndex = (uint16_t)(m_pipelines.size() - 1);
    rec.downsample = downsample;
    rec.maxRepeat = maxRepeat;
    rec.groupW = groupW;
    rec.groupH = groupH;
    m_passes.push_back(rec);
}

Result Graph::create(const InstanceCreationDesc& desc) {
    const LibraryDesc& lib = libraryDesc();
    for (uint32_t i = 0; i < desc.denoisersNum; i++) {
        const DenoiserDesc& dd = desc.denoisers[i];
        bool supported = false;
        for (uint32_t j = 0; j < lib.supportedDenoisersNum; j++) supported |= lib.supportedDenoisers[j] == dd.denoiser;
        if (!supported) return Result::UNSUPPORTED;
        for (uint32_t j = 0; j < desc.denoisersNum; j++)
            if (i != j && desc.denoisers[j].identifier == dd.identifier) return Result::NON_UNIQUE_IDENTIFIER;

        m_permanentBase = (uint16_t)m_permanentPool.size();
        m_transientBase = (uint16_t)m_transientPool.size();
        m_transientRemap.clear();

        DenoiserState d;
        d.desc = dd;
        d.firstPass = m_passes.size();
        d.firstSwap = m_swaps.size();
        size_t firstResource = m_resources.size();

        switch (dd.denoiser) {
            case Denoiser::REBLUR_DIFFUSE: buildReblur(d, true, false); break;
            case Denoiser::REBLUR_SPECULAR: buildReblur(d, false, true); break;
            case Denoiser::REBLUR_DIFFUSE_SPECULAR: buildReblur(d, true, true); break;
            case Denoiser::REBLUR_DIFFUSE_SH: buildReblur(d, true, false, true); break;
            case Denoiser::REBLUR_SPECULAR_SH: buildReblur(d, false, true, true); break;
            case Denoiser::REBLUR_DIFFUSE_SPECULAR_SH: buildReblur(d, true, true, true); break;
            case Denoiser::REBLUR_DIFFUSE_DIRECTIONAL_OCCLUSION: buildReblur(d, true, false, false, true); break;
            case Denoiser::REBLUR_DIFFUSE_OCCLUSION: buildReblurOcclusion(d, true, false); break;
            case Denoiser::REBLUR_SPECULAR_OCCLUSION: buildReblurOcclusion(d, false, true); break;
            case Denoiser::REBLUR_DIFFUSE_SPECULAR_OCCLUSION: buildReblurOcclusion(d, true, true); break;
            case Denoiser::SIGMA_SHADOW: buildSigmaShadow(d, false); break;
            case Denoiser::SIGMA_SHADOW_TRANSLUCENCY: buildSigmaShadow(d, true); break;
            case Denoiser::REFERENCE: buildReference(d); break;
            case Denoiser::RELAX_DIFFUSE: buildRelax(d, true, false, false); break;
            case Denoiser::RELAX_DIFFUSE_SH: buildRelax(d, true, false, true); break;
            case Denoiser::RELAX_SPECULAR: buildRelax(d, false, true, false); break;
            case Denoiser::RELAX_SPECULAR_SH: buildRelax(d, false, true, true); break;
            case Denoiser::RELAX_DIFFUSE_SPECULAR: buildRelax(d, true, true, false); break;
            case Denoiser::RELAX_DIFFUSE_SPECULAR_SH: buildRelax(d, true, true, true); break;
            default: return Result::INVALID_ARGUMENT;
        }
        d.swapNum = m_swaps.size() - d.firstSwap;
        for (size_t p = d.firstPass; p < m_passes.size(); p++) m_passes[p].identifier = dd.identifier;

        // Everything this denoiser binds as a storage texture gets zeroed on CLEAR_AND_RESTART, user textures
        // included, ping and pong both (InstanceImpl.cpp:172-216). OUT_VALIDATION is exempt (may be absent).
        for (size_t r = firstResource; r < m_resources.size(); r++) {
            const ResourceDesc& res = m_resources[r];
            if (res.descriptorType != DescriptorType::STORAGE_TEXTURE || res.type == ResourceType::OUT_VALIDATION) continue;
            bool seen = false;
            for (const ClearTarget& c : m_clears)
                seen |= c.resource.descriptorType == res.descriptorType && c.resource.type == res.type && c.resource.indexInPool == res.indexInPool;
            if (seen) continue;
            bool isInt = false;
            uint16_t ds = 1;
            if (res.type == ResourceType::PERMANENT_POOL || res.type == ResourceType::TRANSIENT_POOL) {
                const TextureDesc& t = res.type == ResourceType::PERMANENT_POOL ? m_permanentPool[res.indexInPool] : m_transientPool[res.indexInPool];
                isInt = isIntegerFormat(t.format);
                ds = t.downsampleFactor;
            }
            m_clears.push_back({dd.identifier, res, ds, isInt});
            for (size_t s = 0; s < d.swapNum; s++) {
                const SwapEntry& sw = m_swaps[d.firstSwap + s];
                if (sw.resourceIndex == r) {
                    m_clears.push_back({dd.identifier, {res.descriptorType, res.type, sw.partner}, ds, isInt});
                    break;
                }
            }
        }
        m_denoisers.push_back(d);
    }

    // Two trailing clear pipelines: float and uint flavours (InstanceImpl.cpp:218-242).
    static const char* kClearNames[2] = {"Clear (f)", "Clear (ui)"};
    for (int i = 0; i < 2; i++) {
        m_clearPass[i] = m_passes.size();
        beginPass(kClearNames[i]);
        out(Slot::user((ResourceType)0));
        emit(i == 0 ? "Clear.cs.hlsl|FLOAT=1" : "Clear.cs.hlsl|FLOAT=0", 16, 16, 0);
    }
    finalize();
    return Result::SUCCESS;
}

void Graph::finalize() {
    m_desc = {};
    m_desc.constantBufferAndSamplersSpaceIndex = 1;  // NRD.hlsli space/register defaults
    m_desc.resourcesSpaceIndex = 0;
    m_desc.constantBufferRegisterIndex = 0;
    m_desc.samplers = kSamplers;
    m_desc.samplersNum = 2;
    m_desc.shaderEntryPoint = "main";
    m_desc.permanentPool = m_permanentPool.data();
    m_desc.permanentPoolSize = (uint32_t)m_permanentPool.size();
    m_desc.transientPool = m_transientPool.data();
    m_desc.transientPoolSize = (uint32_t)m_transientPool.size();

    for (size_t i = 0; i < m_pipelines.size(); i++) m_pipelines[i].resourceRanges = m_ranges.data() + m_pipelineFirstRange[i];
    m_desc.pipelines = m_pipelines.data();
    m_desc.pipelinesNum = (uint32_t)m_pipelines.size();

    // Descriptor-pool limits: one entry per distinct pass name, permutations share it (InstanceImpl.cpp:655-693).
    DescriptorPoolDesc& pool = m_desc.descriptorPoolDesc;
    std::vector<const char*> seen;
    for (const PassRecord& p : m_passes) {
        bool dup = false;
        for (const char* n : seen) dup |= (n == p.name);
        if (dup) continue;
        seen.push_back(p.name);
        pool.setsMaxNum += p.maxRepeat;
        uint32_t tex = 0, storage = 0;
        for (uint32_t i = 0; i < p.resourcesNum; i++) {
            if (m_resources[p.firstResource + i].descriptorType == DescriptorType::TEXTURE) {
                pool.totalTexturesNum += p.maxRepeat;
                tex++;
            } else {
                pool.totalStorageTexturesNum += p.maxRepeat;
                storage++;
            }
        }
        if (p.constantsSize > m_desc.constantBufferMaxDataSize) m_desc.constantBufferMaxDataSize = p.constantsSize;
        if (tex > pool.perSetTexturesMaxNum) pool.perSetTexturesMaxNum = tex;
        if (storage > pool.perSetStorageTexturesMaxNum) pool.perSetStorageTexturesMaxNum = storage;
    }
    pool.setsMaxNum += (uint32_t)m_clears.size();
    pool.totalStorageTexturesNum += (uint32_t)m_clears.size();
}

// ------------------------------------------------------------------------------------------------
// Per frame
// ------------------------------------------------------------------------------------------------
static void loadMat(Mat4& dst, const float* src) { memcpy(dst.m, src, 64); }

static void rotator(float out[4], float angle) {
    float c = std::cos(angle), s = std::sin(angle);
    out[0] = c; out[1] = s; out[2] = -s; out[3] = c;
}

// Sequence::Weyl1D (ml.hlsli:1712): frac(p + float(n * 10368889) / exp2(24)), uint32 wrap-around on the product.
// Compiled as C++ the reference evaluates exp2() and therefore the sum and frac() in double (tests/test_oracle_math.py
// pins this against the reference MathLib), which matters once n * 10368889 / 2^24 reaches the hundreds.
static float weyl1D(float p, uint32_t n) {
    double x = (double)p + (double)float(n * 10368889u) / 16777216.0;
    return (float)(x - std::floor(x));
}

Result Graph::setCommonSettings(const CommonSettings& cs) {
    m_frame.splitScreenPrev = m_common.splitScreen;
    m_common = cs;

    if (m_firstUse) {
        m_common.accumulationMode = AccumulationMode::CLEAR_AND_RESTART;
        m_firstUse = false;
    }
    if (m_common.accumulationMode != AccumulationMode::CONTINUE) {
        // History is dropped: "previous" state collapses onto whatever this instance saw last
        m_frame.splitScreenPrev = 0.0f;
        m_frame.worldToViewPrev = m_frame.worldToView;
        m_frame.viewToClipPrev = m_frame.viewToClip;
        for (int i = 0; i < 2; i++) {
            m_common.resourceSizePrev[i] = m_common.resourceSize[i];
            m_common.rectSizePrev[i] = m_common.rectSize[i];
            m_common.cameraJitterPrev[i] = m_common.cameraJitter[i];
        }
    }

    const CommonSettings& c = m_common;
    bool ok = c.viewZScale > 0.0f;
    ok &= c.resourceSize[0] != 0 && c.resourceSize[1] != 0 && c.resourceSizePrev[0] != 0 && c.resourceSizePrev[1] != 0;
    ok &= c.rectSize[0] != 0 && c.rectSize[1] != 0 && c.rectSizePrev[0] != 0 && c.rectSizePrev[1] != 0;
    ok &= (c.motionVectorScale[0] != 0.0f && c.motionVectorScale[1] != 0.0f) || c.isMotionVectorInWorldSpace;
    for (int i = 0; i < 2; i++) {
        ok &= c.cameraJitter[i] >= -0.5f && c.cameraJitter[i] <= 0.5f;
        ok &= c.cameraJitterPrev[i] >= -0.5f && c.cameraJitterPrev[i] <= 0.5f;
    }
    ok &= c.denoisingRange > 0.0f && c.disocclusionThreshold > 0.0f && c.disocclusionThresholdAlternate > 0.0f;
    ok &= c.rectOrigin[0] == 0 && c.rectOrigin[1] == 0;  // NRD_SUPPORTS_VIEWPORT_OFFSET = 0 build

    {
        float anglePre = weyl1D(1.0f / std::sqrt(2.0f), c.frameIndex);
        float angle = weyl1D(1.0f / std::sqrt(3.0f), c.frameIndex);
        const float deg = 3.14159265358979323846f / 180.0f;
        rotator(m_frame.rotatorPre, anglePre * (90.0f * deg));
        rotator(m_frame.rotator, angle * (90.0f * deg));
        rotator(m_frame.rotatorPost, angle * (90.0f * deg) + 22.5f * deg);
    }

    FrameState& f = m_frame;
    loadMat(f.viewToClip, c.viewToClipMatrix);
    loadMat(f.viewToClipPrev, c.viewToClipMatrixPrev);
    loadMat(f.worldToView, c.worldToViewMatrix);
    loadMat(f.worldToViewPrev, c.worldToViewMatrixPrev);
    loadMat(f.worldPrevToWorld, c.worldPrevToWorldMatrix);
    // NB: user-provided prev matrices are always the ones used, reset or not (InstanceImpl.cpp:262-267 is
    // overwritten by :350-366).

    mat::Projection proj = mat::decomposeProjection(f.viewToClip);
    memcpy(f.frustum, proj.frustum, 16);
    if (!proj.leftHanded) {
        // Right-handed input: flip view-space Z everywhere so kernels only ever see LH
        mat::negateColumn(f.viewToClip, 2);
        mat::negateColumn(f.viewToClipPrev, 2);
        mat::negateRow(f.worldToView, 2);
        mat::negateRow(f.worldToViewPrev, 2);
    }

    f.viewToWorld = mat::invertOrtho(f.worldToView);
    f.viewToWorldPrev = mat::invertOrtho(f.worldToViewPrev);
    float delta[3];
    for (int i = 0; i < 3; i++) delta[i] = f.viewToWorldPrev.m[12 + i] - f.viewToWorld.m[12 + i];

    // Camera-relative matrices: current camera sits at the origin, previous at "delta"
    mat::setTranslation(f.viewToWorld, 0, 0, 0);
    f.worldToView = mat::invertOrtho(f.viewToWorld);
    mat::setTranslation(f.viewToWorldPrev, delta[0], delta[1], delta[2]);
    f.worldToViewPrev = mat::invertOrtho(f.viewToWorldPrev);

    f.worldToClip = mat::mul(f.viewToClip, f.worldToView);
    f.worldToClipPrev = mat::mul(f.viewToClipPrev, f.worldToViewPrev);

    proj = mat::decomposeProjection(f.viewToClip);
    memcpy(f.frustum, proj.frustum, 16);
    f.projectY = proj.projectY;
    f.orthoMode = proj.ortho ? -1.0f : 0.0f;
    mat::Projection projPrev = mat::decomposeProjection(f.viewToClipPrev);
    memcpy(f.frustumPrev, projPrev.frustum, 16);

    for (int i = 0; i < 3; i++) {
        f.viewDirection[i] = -f.viewToWorld.m[8 + i];
        f.viewDirectionPrev[i] = -f.viewToWorldPrev.m[8 + i];
        f.cameraDelta[i] = delta[i];
    }

    // Frame time: user value if given, else a smoothed wall-clock delta (Timer.cpp:50-59)
    double now = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
    if (m_lastTimeMs >= 0.0) {
        float d = (float)(now - m_lastTimeMs);
        float rel = std::fabs(d - m_smoothedDeltaMs) / (std::fmin(d, m_smoothedDeltaMs) + 1e-7f);
        float k = rel / (1.0f + rel);
        m_smoothedDeltaMs += (d - m_smoothedDeltaMs) * std::fmax(k, 1.0f / 32.0f);
    }
    m_lastTimeMs = now;
    f.timeDelta = c.timeDeltaBetweenFrames > 0.0f ? c.timeDeltaBetweenFrames : m_smoothedDeltaMs;
    f.frameRateScale = std::fmax(33.333f / f.timeDelta, 1.0f);

    float dx = std::fabs(c.cameraJitter[0] - c.cameraJitterPrev[0]);
    float dy = std::fabs(c.cameraJitter[1] - c.cameraJitterPrev[1]);
    f.jitterDelta = std::fmax(dx, dy);
    float fps = f.frameRateScale * 30.0f;
    float nonLinearAccumSpeed = fps * 0.25f / (1.0f + fps * 0.25f);
    f.checkerboardResolveAccumSpeed = nonLinearAccumSpeed + (0.5f - nonLinearAccumSpeed) * f.jitterDelta;

    return ok ? Result::SUCCESS : Result::INVALID_ARGUMENT;
}

Result Graph::setDenoiserSettings(Identifier id, const void* settings) {
    for (DenoiserState& d : m_denoisers) {
        if (d.desc.identifier != id) continue;
        memcpy(&d.settings, settings, d.settingsSize);
        return Result::SUCCESS;
    }
    return Result::INVALID_ARGUMENT;
}

void Graph::flipPingPong(const DenoiserState& d) {
    for (size_t i = 0; i < d.swapNum; i++) {
        SwapEntry& s = m_swaps[d.firstSwap + i];
        std::swap(m_resources[s.resourceIndex].indexInPool, s.partner);
    }
}

void* Graph::pushDispatch(const DenoiserState& d, uint32_t localPassIndex) {
    const PassRecord& p = m_passes[d.firstPass + localPassIndex];
    DispatchDesc out = {};
    out.name = p.name;
    out.identifier = p.identifier;
    out.resources = m_resources.data() + p.firstResource;
    out.resourcesNum = p.resourcesNum;
    out.pipelineIndex = p.pipelineIndex;
    out.constantBufferDataSize = p.constantsSize;
    if (m_constantOffset + p.constantsSize <= kConstantArenaSize) {
        out.constantBufferData = m_constantData + m_constantOffset;
        memset((void*)out.constantBufferData, 0, p.constantsSize);
    }
    m_constantOffset += p.constantsSize;

    uint16_t w = m_common.rectSize[0], h = m_common.rectSize[1], ds = p.downsample;
    if (ds == GRID_FROM_PREV_RECT) {
        w = m_common.rectSizePrev[0];
        h = m_common.rectSizePrev[1];
        ds = 1;
    } else if (ds == GRID_FROM_RESOURCE) {
        w = m_common.resourceSize[0];
        h = m_common.resourceSize[1];
        ds = 1;
    }
    out.gridWidth = divUp(divUp(w, ds), p.groupW);
    out.gridHeight = divUp(divUp(h, ds), p.groupH);
    m_active.push_back(out);
    return (void*)out.constantBufferData;
}

Result Graph::getComputeDispatches(const Identifier* ids, uint32_t idsNum, const DispatchDesc*& out, uint32_t& outNum) {
    m_constantOffset = 0;
    m_active.clear();
    if (!ids || !idsNum) {
        out = nullptr;
        outNum = 0;
        return !idsNum ? Result::SUCCESS : Result::INVALID_ARGUMENT;
    }

    if (m_common.accumulationMode == AccumulationMode::CLEAR_AND_RESTART) {
        for (const ClearTarget& c : m_clears) {
            if (!contains(ids, idsNum, c.identifier)) continue;
            const PassRecord& p = m_passes[m_clearPass[c.isInteger ? 1 : 0]];
            DispatchDesc d = {};
            d.name = p.name;
            d.identifier = c.identifier;
            d.resources = &c.resource;
            d.resourcesNum = 1;
            d.pipelineIndex = p.pipelineIndex;
            d.gridWidth = divUp(divUp(m_common.resourceSize[0], c.downsample), p.groupW);
            d.gridHeight = divUp(divUp(m_common.resourceSize[1], c.downsample), p.groupH);
            m_active.push_back(d);
        }
    }

    for (const DenoiserState& d : m_denoisers) {
        if (!contains(ids, idsNum, d.desc.identifier)) continue;
        flipPingPong(d);
        switch (d.desc.denoiser) {
            case Denoiser::REBLUR_DIFFUSE:
            case Denoiser::REBLUR_SPECULAR:
            case Denoiser::REBLUR_DIFFUSE_SH:
            case Denoiser::REBLUR_SPECULAR_SH:
            case Denoiser::REBLUR_DIFFUSE_SPECULAR_SH:
            case Denoiser::REBLUR_DIFFUSE_DIRECTIONAL_OCCLUSION:
            case Denoiser::REBLUR_DIFFUSE_SPECULAR: updateReblur(d); break;
            case Denoiser::REBLUR_DIFFUSE_OCCLUSION:
            case Denoiser::REBLUR_SPECULAR_OCCLUSION:
            case Denoiser::REBLUR_DIFFUSE_SPECULAR_OCCLUSION: updateReblurOcclusion(d); break;
            case Denoiser::SIGMA_SHADOW:
            case Denoiser::SIGMA_SHADOW_TRANSLUCENCY: updateSigma(d); break;
            case Denoiser::REFERENCE: updateReference(d); break;
            case Denoiser::RELAX_DIFFUSE:
            case Denoiser::RELAX_DIFFUSE_SH:
            case Denoiser::RELAX_SPECULAR:
            case Denoiser::RELAX_SPECULAR_SH:
            case Denoiser::RELAX_DIFFUSE_SPECULAR_SH:
            case Denoiser::RELAX_DIFFUSE_SPECULAR: updateRelax(d); break;
            default: break;
        }
    }

    for (size_t i = 1; i < m_active.size(); i++) {
        const DispatchDesc& prev = m_active[i - 1];
        DispatchDesc& cur = m_active[i];
        if (prev.constantBufferDataSize == cur.constantBufferDataSize && cur.constantBufferData && prev.constantBufferData &&
            !memcmp(prev.constantBufferData, cur.constantBufferData, cur.constantBufferDataSize))
            cur.constantBufferDataMatchesPreviousDispatch = true;
        else if (prev.constantBufferDataSize == cur.constantBufferDataSize && cur.constantBufferDataSize == 0)
            cur.constantBufferDataMatchesPreviousDispatch = true;  // memcmp of 0 bytes is "equal" in the reference too
    }

    out = m_active.data();
    outNum = (uint32_t)m_active.size();
    return outNum ? Result::SUCCESS : Result::INVALID_ARGUMENT;
}

}  // namespace nrdb

// ------------------------------------------------------------------------------------------------
// Exported C ABI (External/NRD/Include/NRD.h:60-79, Source/Wrapper.cpp:182-239)
// ------------------------------------------------------------------------------------------------
using namespace nrd;

namespace {

void* NRD_CALL defaultAllocate(void*, size_t size, size_t alignment) {
    void* p = nullptr;
    if (alignment < sizeof(void*)) alignment = sizeof(void*);
    return posix_memalign(&p, alignment, size) == 0 ? p : nullptr;
}
void* NRD_CALL defaultReallocate(void*, void* memory, size_t size, size_t) { return realloc(memory, size); }
void NRD_CALL defaultFree(void*, void* memory) { free(memory); }

struct InstanceBox {
    AllocationCallbacks callbacks;
    nrdb::Graph graph;
};

const char* const kResourceTypeNames[] = {
    "IN_MV", "IN_NORMAL_ROUGHNESS", "IN_VIEWZ", "IN_DIFF_CONFIDENCE", "IN_SPEC_CONFIDENCE", "IN_DISOCCLUSION_THRESHOLD_MIX",
    "IN_DIFF_RADIANCE_HITDIST", "IN_SPEC_RADIANCE_HITDIST", "IN_DIFF_HITDIST", "IN_SPEC_HITDIST", "IN_DIFF_DIRECTION_HITDIST",
    "IN_DIFF_SH0", "IN_DIFF_SH1", "IN_SPEC_SH0", "IN_SPEC_SH1", "IN_PENUMBRA", "IN_TRANSLUCENCY", "IN_SIGNAL",
    "OUT_DIFF_RADIANCE_HITDIST", "OUT_SPEC_RADIANCE_HITDIST", "OUT_DIFF_SH0", "OUT_DIFF_SH1", "OUT_SPEC_SH0", "OUT_SPEC_SH1",
    "OUT_DIFF_HITDIST", "OUT_SPEC_HITDIST", "OUT_DIFF_DIRECTION_HITDIST", "OUT_SHADOW_TRANSLUCENCY", "OUT_SIGNAL", "OUT_VALIDATION",
    "TRANSIENT_POOL", "PERMANENT_POOL",
};
static_assert(sizeof(kResourceTypeNames) / sizeof(char*) == (size_t)ResourceType::MAX_NUM, "one name per ResourceType");

const char* const kDenoiserNames[] = {
    "REBLUR_DIFFUSE", "REBLUR_DIFFUSE_OCCLUSION", "REBLUR_DIFFUSE_SH", "REBLUR_SPECULAR", "REBLUR_SPECULAR_OCCLUSION",
    "REBLUR_SPECULAR_SH", "REBLUR_DIFFUSE_SPECULAR", "REBLUR_DIFFUSE_SPECULAR_OCCLUSION", "REBLUR_DIFFUSE_SPECULAR_SH",
    "REBLUR_DIFFUSE_DIRECTIONAL_OCCLUSION", "RELAX_DIFFUSE", "RELAX_DIFFUSE_SH", "RELAX_SPECULAR", "RELAX_SPECULAR_SH",
    "RELAX_DIFFUSE_SPECULAR", "RELAX_DIFFUSE_SPECULAR_SH", "SIGMA_SHADOW", "SIGMA_SHADOW_TRANSLUCENCY", "REFERENCE",
};
static_assert(sizeof(kDenoiserNames) / sizeof(char*) == (size_t)Denoiser::MAX_NUM, "one name per Denoiser");

}  // namespace

#define NRD_EXPORT extern "C" __attribute__((visibility("default")))

NRD_EXPORT Result NRD_CALL CreateInstance(const InstanceCreationDesc& desc, Instance*& instance) {
    AllocationCallbacks cb = desc.allocationCallbacks;
    if (!cb.Allocate || !cb.Reallocate || !cb.Free) cb = {defaultAllocate, defaultReallocate, defaultFree, nullptr};
    void* mem = cb.Allocate(cb.userArg, sizeof(InstanceBox), alignof(InstanceBox));
    if (!mem) return Result::FAILURE;
    InstanceBox* box = new (mem) InstanceBox();
    box->callbacks = cb;
    Result r = box->graph.create(desc);
    if (r != Result::SUCCESS) {
        box->~InstanceBox();
        cb.Free(cb.userArg, mem);
        instance = nullptr;
        return r;
    }
    instance = (Instance*)box;
    return Result::SUCCESS;
}

NRD_EXPORT void NRD_CALL DestroyInstance(Instance& instance) {
    InstanceBox* box = (InstanceBox*)&instance;
    AllocationCallbacks cb = box->callbacks;
    box->~InstanceBox();
    cb.Free(cb.userArg, box);
}

NRD_EXPORT const LibraryDesc* NRD_CALL GetLibraryDesc() { return &nrdb::libraryDesc(); }

NRD_EXPORT const InstanceDesc* NRD_CALL GetInstanceDesc(const Instance& instance) { return &((const InstanceBox*)&instance)->graph.desc(); }

NRD_EXPORT Result NRD_CALL SetCommonSettings(Instance& instance, const CommonSettings& cs) {
    return ((InstanceBox*)&instance)->graph.setCommonSettings(cs);
}

NRD_EXPORT Result NRD_CALL SetDenoiserSettings(Instance& instance, Identifier id, const void* settings) {
    return ((InstanceBox*)&instance)->graph.setDenoiserSettings(id, settings);
}

NRD_EXPORT Result NRD_CALL GetComputeDispatches(Instance& instance, const Identifier* ids, uint32_t idsNum, const DispatchDesc*& out,
                                                uint32_t& outNum) {
    return ((InstanceBox*)&instance)->graph.getComputeDispatches(ids, idsNum, out, outNum);
}

// NB: the reference's name table is out of step with its enum (Wrapper.cpp:53-88 lists the confidence inputs after
// the SH inputs), so it mislabels most inputs. This library returns the name of the enumerator that was passed.
NRD_EXPORT const char* GetResourceTypeString(ResourceType t) {
    uint32_t i = (uint32_t)t;
    return i < (uint32_t)ResourceType::MAX_NUM ? kResourceTypeNames[i] : nullptr;
}

NRD_EXPORT const char* GetDenoiserString(Denoiser d) {
    uint32_t i = (uint32_t)d;
    return i < (uint32_t)Denoiser::MAX_NUM ? kDenoiserNames[i] : nullptr;
}
