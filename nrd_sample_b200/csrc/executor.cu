// CUDA executor for NRD dispatch streams (include/nrdcu.h). Host code only — the kernels live in kernels/*.cu.
// Mirrors the reference's executor semantics (External/NRD/Integration/NRDIntegration.hpp): pools sized
// resource / downsampleFactor (:246-318), bindings resolved per dispatch (:757-766), dispatches issued in order on one
// queue (here: one CUDA stream, which gives the same "barrier between every pair of dependent dispatches").
#include <cuda_runtime.h>
#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/nrd_b200.h"
#include "../../include/nrdcu.h"
#include "pipeline_key.h"
#include "kernels/reblur_common.cuh"
#include "kernels/peer_halo.cuh"
#include "kernels/sigma_common.cuh"

namespace nrdk {
// kernels/*.cu
void launchClear(void* data, int rowBytes, int height, int pitch, cudaStream_t stream);
void launchReblurClassifyTiles(const ReblurConstants&, const ClassifyTilesParams&, Rows, cudaStream_t);
void launchReblurValidation(const ReblurConstants&, const ReblurValidationParams&, cudaStream_t);
void launchReblurGeometryPlane(const ReblurConstants&, const GeometryPlaneParams&, int row0, int row1, cudaStream_t);
void launchReblurHitDistReconstruction(const ReblurConstants&, const HitDistReconstructionParams&, int signal, bool occlusion, bool is5x5, Rows, cudaStream_t);
void launchReblurPrePass(const ReblurConstants&, const PrePassParams&, int signal, int flags, Rows, cudaStream_t);
void launchReblurSplitScreen(const ReblurConstants&, const SplitScreenParams&, int signal, Rows, cudaStream_t);
void launchReblurBlur(const ReblurConstants&, const BlurParams&, int signal, int flags, Rows, cudaStream_t);
void launchReblurPostBlur(const ReblurConstants&, const PostBlurParams&, int signal, bool temporalStabilization, int flags, Rows, cudaStream_t);
void launchReblurTemporalAccumulation(const ReblurConstants&, const TemporalAccumulationParams&, int signal, int mode, Rows, cudaStream_t);
void launchReblurHistoryFix(const ReblurConstants&, const HistoryFixParams&, int signal, int mode, bool quads, Rows, cudaStream_t);
void launchReblurTemporalStabilization(const ReblurConstants&, const TemporalStabilizationParams&, int signal, int mode, Rows, cudaStream_t);
bool readMirrorProbe(unsigned long long* out, bool reset);
void sigmaSetCopyFusion(bool on);
void sigmaFlushPendingCopy(cudaStream_t stream);
uint32_t dispatchSigma(const PipelineKey& key, const void* constants, uint32_t constantsSize, const nrdcuTexture* t, uint32_t n, Rows rows, cudaStream_t stream, std::string& err);
uint32_t dispatchRelax(const PipelineKey& key, const void* constants, uint32_t constantsSize, const nrdcuTexture* t, uint32_t n, Rows rows, cudaStream_t stream, std::string& err);
uint32_t dispatchReference(const PipelineKey& key, const void* constants, uint32_t constantsSize, const nrdcuTexture* t, uint32_t n, uint32_t gridW, uint32_t gridH, cudaStream_t stream,
                           std::string& err);
}  // namespace nrdk

using namespace nrd;

namespace {

thread_local std::string g_lastError;
std::atomic<uint64_t> g_launchCount{0};

// The geometry plane of the frame being denoised ( kernels/reblur_spatial.cu: { normal, |viewZ| } decoded once per texel for the 48 spatial taps per
// pixel ). nrdcuDenoiseRows points g_plane at its context's plane for the duration of the call; the context-less nrdcuDispatch builds a
// stream-ordered temporary instead.
struct GeomPlane {
    nrdcuTexture tex = {};
    const void *fromNormalRoughness = nullptr, *fromViewZ = nullptr;   // what it was decoded from ...
    const void* viewZCopy = nullptr;                                    // ... and the copy of that viewZ the blur pass wrote ( PREV_VIEWZ: what post-blur binds )
    int row0 = 0, row1 = 0;                                             // ... and for which rows; row1 == 0: not decoded this frame
    uint32_t margin = 64;                                               // rows beyond the strip the taps may reach ( strips only )
    int stripRow0 = 0, stripRow1 = 0x7FFFFFFF;                          // the strip this context computes ( a pass may arrive as several row ranges of it )
};
thread_local GeomPlane* g_plane = nullptr;

uint32_t fail(Result r, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_lastError = buf;
    return (uint32_t)r;
}

uint32_t bytesPerTexel(uint32_t fmt) {
    switch ((Format)fmt) {
        case Format::R8_UNORM: case Format::R8_UINT: return 1;
        case Format::RG8_UNORM: case Format::R16_UINT: case Format::R16_SFLOAT: case Format::R16_UNORM: return 2;
        case Format::RGBA8_UNORM: case Format::RG16_SFLOAT: case Format::R32_UINT: case Format::R32_SFLOAT: case Format::R10_G10_B10_A2_UNORM: return 4;
        case Format::RGBA16_SFLOAT: case Format::RGBA16_SNORM: case Format::RGBA16_UNORM: return 8;
        case Format::RGBA32_SFLOAT: return 16;
        default: return 0;
    }
}

// Typed view construction with format checking
struct Binder {
    const nrdcuTexture* t;
    uint32_t n, next = 0;
    bool ok = true;
    std::string* err;
    const char* shader;
    template <class V> V take(Format expect) {
        V v{};
        if (next >= n) {
            ok = false;
            return v;
        }
        const nrdcuTexture& x = t[next];
        if (x.format != (uint32_t)expect || !x.data || (x.pitchBytes % bytesPerTexel(x.format)) != 0 || x.pitchBytes < x.width * bytesPerTexel(x.format)) {
            if (ok) {
                char buf[256];
                snprintf(buf, sizeof(buf), "%s: binding %u has format %u pitch %u (expected format %u)", shader, next, x.format, x.pitchBytes, (uint32_t)expect);
                *err = buf;
            }
            ok = false;
        }
        v.data = (uint8_t*)x.data;
        v.w = (int)x.width;
        v.h = (int)x.height;
        v.pitch = (int)(x.pitchBytes / (bytesPerTexel(x.format) ? bytesPerTexel(x.format) : 1u));  // in texels (see TexView)
        v.fmt = x.format;
        next++;
        return v;
    }
    // NRD_MODE = OCCLUSION / DO: the lobe's signal ( and fast history ) lives in 16 / 8-bit UNORM / SNORM pool textures, while the application may hand over
    // anything from R8 to RGBA16F for IN / OUT_*_HITDIST ( NRDSample binds its RGBA16F textures, Source/NRDSample.cpp:489-500 ): the kernels read the format
    // from the view ( reblur_common.cuh Sig< MODE > ), so any of the listed formats is accepted here
    template <class V> V takeAny(const std::initializer_list<Format>& accepted) {
        V v{};
        if (next >= n) {
            ok = false;
            return v;
        }
        const nrdcuTexture& x = t[next];
        bool known = false;
        for (Format f : accepted) known |= x.format == (uint32_t)f;
        const uint32_t bpp = bytesPerTexel(x.format);
        if (!known || !bpp || !x.data || (x.pitchBytes % bpp) != 0 || x.pitchBytes < x.width * bpp) {
            if (ok) {
                char buf[256];
                snprintf(buf, sizeof(buf), "%s: binding %u has format %u pitch %u, which this mode cannot read", shader, next, x.format, x.pitchBytes);
                *err = buf;
            }
            ok = false;
        }
        v.data = (uint8_t*)x.data;
        v.w = (int)x.width;
        v.h = (int)x.height;
        v.pitch = (int)(x.pitchBytes / (bpp ? bpp : 1u));
        v.fmt = x.format;
        next++;
        return v;
    }
    // optional single-channel guides: IN_VIEWZ as a dummy when disabled, otherwise whatever the application owns (any size, see guideFormat)
    nrdk::TexAnyX takeGuide() {
        nrdk::TexAnyX v{};
        if (next >= n) {
            ok = false;
            return v;
        }
        const nrdcuTexture& x = t[next];
        if (!nrdk::bindGuide(x.format, x.data, x.width, x.height, x.pitchBytes, v)) {
            if (ok) {
                char buf[256];
                snprintf(buf, sizeof(buf), "%s: binding %u (single-channel guide) has unsupported format %u / pitch %u", shader, next, x.format, x.pitchBytes);
                *err = buf;
            }
            ok = false;
        }
        next++;
        return v;
    }
};

bool checkLaunch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        g_lastError = std::string(what) + ": " + cudaGetErrorString(e);
        return false;
    }
    g_launchCount.fetch_add(1, std::memory_order_relaxed);
    return true;
}

uint32_t dispatchReblur(const nrdk::PipelineKey& key, const void* constants, uint32_t constantsSize, const nrdcuTexture* tex, uint32_t n, uint32_t flags, nrdk::Rows rows,
                        cudaStream_t stream) {
    using namespace nrdk;
    const char* const id = key.id;
    if (constantsSize != sizeof(ReblurConstants) || !constants) return fail(Result::INVALID_ARGUMENT, "%s: expected %zu constant bytes, got %u", id, sizeof(ReblurConstants), constantsSize);
    ReblurConstants cb;
    memcpy(&cb, constants, sizeof(cb));
    // dynamic resolution ( rectSize < resourceSize, CommonSettings ): the kernels work in rect pixels and scale uv by gResolutionScale( Prev ) wherever the
    // reference samples a resource-sized texture; rectOrigin is always 0 ( NRD_SUPPORTS_VIEWPORT_OFFSET = 0: the host library rejects anything else )
    if ((cb.diffCheckerboard == 2) != (cb.specCheckerboard == 2)) return fail(Result::INVALID_ARGUMENT, "%s: checkerboard constants %u / %u", id, cb.diffCheckerboard, cb.specCheckerboard);

    const bool quads = flags & NRDCU_FLAG_QUAD_INTRINSICS;
    const int kflags = (quads ? 1 : 0) | ((flags & NRDCU_FLAG_ROBUST_MIRROR_TEST) ? 2 : 0) | ((flags & NRDCU_FLAG_PROBE_MIRROR) ? 4 : 0);
    std::string err;
    Binder b{tex, n, 0, true, &err, id};
    // "<file>|NRD_SIGNAL=<DIFF|SPEC|BOTH>|NRD_MODE=<RADIANCE|SH|OCCLUSION|DO><suffix>" (InstanceImpl.h:59-67), parsed into `key` when the pipeline was
    // resolved. REBLUR_DIFFUSE / REBLUR_SPECULAR bind only their own lobe's textures (REBLUR_*.resources.hlsli): takeD / takeS consume a binding only when
    // the permutation has that lobe. NRD_MODE=SH ( REBLUR_DIFFUSE_SH / REBLUR_SPECULAR_SH / REBLUR_DIFFUSE_SPECULAR_SH ): each lobe binds a second RGBA16F
    // next to the first ( takeShD / takeShS below, in the order of the REBLUR_*.resources.hlsli lists )
    const int signal = key.signal, mode = key.mode;
    const bool sh = mode == MODE_SH, occlusion = mode == MODE_OCCLUSION, polymorphicMode = mode == MODE_OCCLUSION || mode == MODE_DO;
    const int kflagsMode = kflags | (mode << 4);
    const bool hasDiff = (signal & 1) != 0, hasSpec = (signal & 2) != 0;
    const uint32_t lobes = (hasDiff ? 1u : 0u) + (hasSpec ? 1u : 0u);
    auto done = [&](uint32_t expected) -> uint32_t {
        if (!b.ok || b.next != expected || n != expected) return fail(Result::INVALID_ARGUMENT, "%s", err.empty() ? (std::string(id) + ": wrong number of textures").c_str() : err.c_str());
        return 0xFFFFFFFFu;
    };
    // the lobe's signal texture: RGBA16F in the RADIANCE / SH modes; anything the format-polymorphic accessors read in the occlusion modes and in the two
    // RADIANCE permutations the occlusion denoisers share ( hit-distance reconstruction of DO, split screen ): R16_UNORM / RGBA16_SNORM pools, application formats
    bool anySignalFormat = polymorphicMode;
    auto takeSig = [&]() {
        if (!anySignalFormat) return b.take<TexRGBA16F>(Format::RGBA16_SFLOAT);
        return b.takeAny<TexRGBA16F>({Format::R8_UNORM, Format::R16_UNORM, Format::R16_SFLOAT, Format::R32_SFLOAT, Format::RGBA8_UNORM, Format::RGBA16_UNORM, Format::RGBA16_SNORM,
                                      Format::RGBA16_SFLOAT, Format::RGBA32_SFLOAT});
    };
    auto takeD16 = [&]() { return hasDiff ? takeSig() : TexRGBA16F{}; };
    auto takeS16 = [&]() { return hasSpec ? takeSig() : TexRGBA16F{}; };
    auto takeShD = [&]() { return (hasDiff && sh) ? b.take<TexRGBA16F>(Format::RGBA16_SFLOAT) : TexRGBA16F{}; };
    auto takeShS = [&]() { return (hasSpec && sh) ? b.take<TexRGBA16F>(Format::RGBA16_SFLOAT) : TexRGBA16F{}; };
    const uint32_t shLobes = sh ? lobes : 0u;
    // the sky-tile mask: R8_UNORM, or the full-resolution RGBA16F that REBLUR_DIFFUSE_SPECULAR_SH's pool table puts under TILES ( see TexTiles )
    auto takeTiles = [&]() {
        TexTiles v{};
        if (b.next < n && tex[b.next].format == (uint32_t)Format::RGBA16_SFLOAT) {
            static_cast<TexView&>(v) = b.take<TexRGBA16F>(Format::RGBA16_SFLOAT);
            v.texelBytes = 8u;
        } else {
            static_cast<TexView&>(v) = b.take<TexR8>(Format::R8_UNORM);
            v.texelBytes = 1u;
        }
        return v;
    };
    auto takeFast = [&]() { return polymorphicMode ? b.takeAny<TexR16F>({Format::R8_UNORM, Format::R16_UNORM, Format::R16_SFLOAT}) : b.take<TexR16F>(Format::R16_SFLOAT); };
    auto takeDF = [&]() { return hasDiff ? takeFast() : TexR16F{}; };                                             // fast history ( REBLUR_FAST_TYPE )
    auto takeSF = [&]() { return hasSpec ? b.take<TexR16F>(Format::R16_SFLOAT) : TexR16F{}; };                    // hit distance for tracking: always R16F
    auto takeSFast = [&]() { return hasSpec ? takeFast() : TexR16F{}; };
    // data1: RG8_UNORM for two lobes, R8_UNORM for one (Reblur.cpp: DATA1 format); P has `data1` + `data1R8` or `outData1` + `outData1R8`
    auto takeData1 = [&](TexRG8& both, TexR8& single) {
        if (lobes == 2) both = b.take<TexRG8>(Format::RG8_UNORM);
        else single = b.take<TexR8>(Format::R8_UNORM);
    };

    // the geometry plane for a spatial pass reading ( normalRoughness, viewZ ): decoded at most once per frame and row range
    void* tempPlane = nullptr;
    auto acquirePlane = [&](const TexNR& nr, const TexR32F& z, TexGeom& out) -> bool {
        const int rectH = cb.rectSizeMinusOne[1] + 1;
        const int begin = g_plane ? std::min(rows.begin, g_plane->stripRow0) : rows.begin, end = g_plane ? std::max(std::min(rows.end, rectH), std::min(g_plane->stripRow1, rectH)) : rows.end;
        const bool strip = begin > 0 || end < rectH;
        const int margin = g_plane ? (int)g_plane->margin : 64;
        const int r0 = strip ? std::max(begin - margin, 0) : 0, r1 = strip ? std::min(end + margin, rectH) : rectH;
        nrdcuTexture t = {};
        bool decode = true;
        if (g_plane) {
            if (!g_plane->tex.data) return false;
            t = g_plane->tex;
            const bool sameZ = g_plane->fromViewZ == z.data || (g_plane->viewZCopy && g_plane->viewZCopy == z.data);
            decode = !(g_plane->fromNormalRoughness == nr.data && sameZ && g_plane->row1 > g_plane->row0 && g_plane->row0 <= r0 && g_plane->row1 >= r1);
            if (decode) {
                g_plane->fromNormalRoughness = nr.data;
                g_plane->fromViewZ = z.data;
                g_plane->viewZCopy = nullptr;
                g_plane->row0 = r0;
                g_plane->row1 = r1;
            }
        } else {
            const uint32_t pitch = ((uint32_t)nr.w * 16u + 255u) & ~255u;
            if (cudaMallocAsync(&tempPlane, (size_t)pitch * nr.h, stream) != cudaSuccess) return false;
            t = {tempPlane, (uint32_t)nr.w, (uint32_t)nr.h, pitch, (uint32_t)Format::RGBA32_SFLOAT};
        }
        out.data = (uint8_t*)t.data;
        out.w = (int)t.width;
        out.h = (int)t.height;
        out.pitch = (int)(t.pitchBytes / 16u);
        if (decode) {
            GeometryPlaneParams gp = {nr, z, out};
            launchReblurGeometryPlane(cb, gp, r0, r1, stream);
            g_launchCount.fetch_add(1, std::memory_order_relaxed);
        }
        return true;
    };
    auto releasePlane = [&]() {
        if (tempPlane) cudaFreeAsync(tempPlane, stream);
    };

    if (key.pass == REBLUR_VALIDATION) {
        // REBLUR_Validation.resources.hlsli:22-34: bound by whatever format each texture has ( data1 RG8 / R8, data2 R32_UINT / R8_UINT or data1 again, the lobe
        // inputs of the denoiser, OUT_VALIDATION "RGBA8+" )
        auto takeView = [&]() {
            TexView v = b.takeAny<TexView>({Format::R8_UNORM, Format::R8_UINT, Format::RG8_UNORM, Format::RGBA8_UNORM, Format::R16_UNORM, Format::R16_SFLOAT, Format::R16_UINT,
                                            Format::RG16_SFLOAT, Format::RGBA16_UNORM, Format::RGBA16_SNORM, Format::RGBA16_SFLOAT, Format::R32_UINT, Format::R32_SFLOAT, Format::RGBA32_SFLOAT});
            return v;
        };
        ReblurValidationParams p = {};
        p.normalRoughness = b.take<TexNR>(Format::R10_G10_B10_A2_UNORM);
        p.viewZ = b.take<TexR32F>(Format::R32_SFLOAT);
        p.mv = takeView();
        p.data1 = takeView();
        p.data2 = takeView();
        p.diff = takeView();
        p.spec = takeView();
        p.out = takeView();
        uint32_t r = done(8);
        if (r != 0xFFFFFFFFu) return r;
        launchReblurValidation(cb, p, stream);
    } else if (key.pass == REBLUR_CLASSIFY_TILES) {
        ClassifyTilesParams p = {};
        p.inViewZ = b.take<TexR32F>(Format::R32_SFLOAT);
        p.outTiles = takeTiles();
        uint32_t r = done(2);
        if (r != 0xFFFFFFFFu) return r;
        launchReblurClassifyTiles(cb, p, rows, stream);
    } else if (key.pass == REBLUR_HITDIST_RECONSTRUCTION) {
        anySignalFormat = true;
        HitDistReconstructionParams p = {};
        p.tiles = takeTiles();
        p.normalRoughness = b.take<TexNR>(Format::R10_G10_B10_A2_UNORM);
        p.viewZ = b.take<TexR32F>(Format::R32_SFLOAT);
        p.inDiff = takeD16();
        p.inSpec = takeS16();
        p.outDiff = takeD16();
        p.outSpec = takeS16();
        uint32_t r = done(3 + 2 * lobes);
        if (r != 0xFFFFFFFFu) return r;
        launchReblurHitDistReconstruction(cb, p, signal, occlusion, key.mode5x5, rows, stream);
    } else if (key.pass == REBLUR_SPLIT_SCREEN) {
        anySignalFormat = true;
        SplitScreenParams p = {};
        p.viewZ = b.take<TexR32F>(Format::R32_SFLOAT);
        p.inDiff = takeD16();
        p.inSpec = takeS16();
        p.inDiffSh = takeShD();
        p.inSpecSh = takeShS();
        p.outDiff = takeD16();
        p.outSpec = takeS16();
        p.outDiffSh = takeShD();
        p.outSpecSh = takeShS();
        uint32_t r = done(1 + 2 * lobes + 2 * shLobes);
        if (r != 0xFFFFFFFFu) return r;
        launchReblurSplitScreen(cb, p, signal, rows, stream);
    } else if (key.pass == REBLUR_PREPASS) {
        PrePassParams p = {};
        p.tiles = takeTiles();
        p.normalRoughness = b.take<TexNR>(Format::R10_G10_B10_A2_UNORM);
        p.viewZ = b.take<TexR32F>(Format::R32_SFLOAT);
        p.inDiff = takeD16();
        p.inSpec = takeS16();
        p.inDiffSh = takeShD();
        p.inSpecSh = takeShS();
        p.outDiff = takeD16();
        p.outSpec = takeS16();
        p.outSpecHitDistForTracking = takeSF();
        p.outDiffSh = takeShD();
        p.outSpecSh = takeShS();
        uint32_t r = done(3 + 2 * lobes + (hasSpec ? 1 : 0) + 2 * shLobes);
        if (r != 0xFFFFFFFFu) return r;
        if (!acquirePlane(p.normalRoughness, p.viewZ, p.geom)) return fail(Result::FAILURE, "%s: no memory for the geometry plane", id);
        launchReblurPrePass(cb, p, signal, kflagsMode, rows, stream);
        releasePlane();
    } else if (key.pass == REBLUR_TEMPORAL_ACCUMULATION) {
        TemporalAccumulationParams p = {};
        p.tiles = takeTiles();
        p.normalRoughness = b.take<TexNR>(Format::R10_G10_B10_A2_UNORM);
        p.viewZ = b.take<TexR32F>(Format::R32_SFLOAT);
        p.mv = b.take<TexRGBA16F>(Format::RGBA16_SFLOAT);
        p.prevViewZ = b.take<TexR32F>(Format::R32_SFLOAT);
        p.prevNormalRoughness = b.take<TexNR>(Format::R10_G10_B10_A2_UNORM);
        p.prevInternalData = b.take<TexR16U>(Format::R16_UINT);
        p.disocclusionThresholdMix = b.takeGuide();
        if (hasDiff) p.diffConfidence = b.takeGuide();
        if (hasSpec) p.specConfidence = b.takeGuide();
        p.inDiff = takeD16();
        p.inSpec = takeS16();
        p.historyDiff = takeD16();
        p.historySpec = takeS16();
        p.historyDiffFast = takeDF();
        p.historySpecFast = takeSFast();
        p.prevSpecHitDistForTracking = takeSF();
        if (!occlusion) p.inSpecHitDistForTracking = takeSF();   // no pre-pass in front of the occlusion denoisers ( REBLUR_TemporalAccumulation.resources.hlsli:39-41 )
        p.inDiffSh = takeShD();
        p.inSpecSh = takeShS();
        p.historyDiffSh = takeShD();
        p.historySpecSh = takeShS();
        takeData1(p.outData1, p.outData1R8);
        p.outDiff = takeD16();
        p.outSpec = takeS16();
        p.outDiffFast = takeDF();
        p.outSpecFast = takeSFast();
        p.outSpecHitDistForTracking = takeSF();
        if (!occlusion) {   // the occlusion denoisers write no data2 ( :63-65 )
            if (hasSpec) p.outData2 = b.take<TexR32U>(Format::R32_UINT);
            else p.outData2R8 = b.take<TexR8U>(Format::R8_UINT);
        }
        p.outDiffSh = takeShD();
        p.outSpecSh = takeShS();
        uint32_t r = done((occlusion ? 9 : 10) + 6 * lobes + (hasSpec ? (occlusion ? 2 : 3) : 0) + 3 * shLobes);
        if (r != 0xFFFFFFFFu) return r;
        launchReblurTemporalAccumulation(cb, p, signal, mode, rows, stream);
    } else if (key.pass == REBLUR_HISTORY_FIX) {
        HistoryFixParams p = {};
        p.tiles = takeTiles();
        p.normalRoughness = b.take<TexNR>(Format::R10_G10_B10_A2_UNORM);
        takeData1(p.data1, p.data1R8);
        p.viewZ = b.take<TexR32F>(Format::R32_SFLOAT);
        p.inDiff = takeD16();
        p.inSpec = takeS16();
        p.inDiffFast = takeDF();
        p.inSpecFast = takeSFast();
        if (!occlusion) p.specHitDistForTracking = takeSF();   // REBLUR_HistoryFix.resources.hlsli:30-32
        p.inDiffSh = takeShD();
        p.inSpecSh = takeShS();
        p.outDiff = takeD16();
        p.outSpec = takeS16();
        p.outDiffFast = takeDF();
        p.outSpecFast = takeSFast();
        p.outDiffSh = takeShD();
        p.outSpecSh = takeShS();
        uint32_t r = done(4 + 4 * lobes + (hasSpec && !occlusion ? 1 : 0) + 2 * shLobes);
        if (r != 0xFFFFFFFFu) return r;
        launchReblurHistoryFix(cb, p, signal, mode, quads, rows, stream);
    } else if (key.pass == REBLUR_BLUR) {
        BlurParams p = {};
        p.tiles = takeTiles();
        p.normalRoughness = b.take<TexNR>(Format::R10_G10_B10_A2_UNORM);
        p.viewZ = b.take<TexR32F>(Format::R32_SFLOAT);
        takeData1(p.data1, p.data1R8);
        p.inDiff = takeD16();
        p.inSpec = takeS16();
        p.inDiffSh = takeShD();
        p.inSpecSh = takeShS();
        p.outViewZ = b.take<TexR32F>(Format::R32_SFLOAT);
        p.outDiff = takeD16();
        p.outSpec = takeS16();
        p.outDiffSh = takeShD();
        p.outSpecSh = takeShS();
        uint32_t r = done(5 + 2 * lobes + 2 * shLobes);
        if (r != 0xFFFFFFFFu) return r;
        if (!acquirePlane(p.normalRoughness, p.viewZ, p.geom)) return fail(Result::FAILURE, "%s: no memory for the geometry plane", id);
        launchReblurBlur(cb, p, signal, kflagsMode, rows, stream);
        releasePlane();
        // the blur pass copies viewZ ( sky included ) into PREV_VIEWZ, which is what post-blur binds as its viewZ: same texels, same plane
        if (g_plane && g_plane->fromViewZ == p.viewZ.data) g_plane->viewZCopy = p.outViewZ.data;
    } else if (key.pass == REBLUR_POST_BLUR) {
        const bool ts = key.temporalStabilization;
        PostBlurParams p = {};
        p.tiles = takeTiles();
        p.normalRoughness = b.take<TexNR>(Format::R10_G10_B10_A2_UNORM);
        takeData1(p.data1, p.data1R8);
        p.viewZ = b.take<TexR32F>(Format::R32_SFLOAT);
        p.inDiff = takeD16();
        p.inSpec = takeS16();
        p.inDiffSh = takeShD();
        p.inSpecSh = takeShS();
        p.outNormalRoughness = b.take<TexNR>(Format::R10_G10_B10_A2_UNORM);
        p.outDiff = takeD16();
        p.outSpec = takeS16();
        if (!ts) {
            p.outInternalData = b.take<TexR16U>(Format::R16_UINT);
            p.outDiffCopy = takeD16();
            p.outSpecCopy = takeS16();
            p.outDiffShCopy = takeShD();
            p.outSpecShCopy = takeShS();
        }
        p.outDiffSh = takeShD();  // the SH history comes last ( Reblur_DiffuseSpecularSh.hpp:268-281 )
        p.outSpecSh = takeShS();
        uint32_t r = done(ts ? 5 + 2 * lobes + 2 * shLobes : 6 + 3 * lobes + 3 * shLobes);
        if (r != 0xFFFFFFFFu) return r;
        if (!acquirePlane(p.normalRoughness, p.viewZ, p.geom)) return fail(Result::FAILURE, "%s: no memory for the geometry plane", id);
        launchReblurPostBlur(cb, p, signal, ts, kflagsMode, rows, stream);
        releasePlane();
    } else if (key.pass == REBLUR_TEMPORAL_STABILIZATION) {
        TemporalStabilizationParams p = {};
        p.tiles = takeTiles();
        p.normalRoughness = b.take<TexNR>(Format::R10_G10_B10_A2_UNORM);
        p.viewZ = b.take<TexR32F>(Format::R32_SFLOAT);
        takeData1(p.data1, p.data1R8);
        if (hasSpec) p.data2 = b.take<TexR32U>(Format::R32_UINT);
        else p.data2R8 = b.take<TexR8U>(Format::R8_UINT);
        p.specHitDistForTracking = takeSF();
        p.inDiff = takeD16();
        p.inSpec = takeS16();
        p.historyDiffLuma = takeDF();
        p.historySpecLuma = takeSF();
        p.inDiffSh = takeShD();  // the SH history written by the post-blur
        p.inSpecSh = takeShS();
        p.mv = b.take<TexRGBA16F>(Format::RGBA16_SFLOAT);
        p.outInternalData = b.take<TexR16U>(Format::R16_UINT);
        p.outDiff = takeD16();
        p.outSpec = takeS16();
        p.outDiffLuma = takeDF();
        p.outSpecLuma = takeSF();
        p.outDiffSh = takeShD();
        p.outSpecSh = takeShS();
        uint32_t r = done(7 + 4 * lobes + (hasSpec ? 1 : 0) + 2 * shLobes);
        if (r != 0xFFFFFFFFu) return r;
        launchReblurTemporalStabilization(cb, p, signal, mode, rows, stream);
    } else {
        return fail(Result::UNSUPPORTED, "no CUDA kernel for shader '%s'", id);
    }
    if (!checkLaunch(id)) return (uint32_t)Result::FAILURE;
    return (uint32_t)Result::SUCCESS;
}

// One dispatch of an already resolved pipeline: no text is looked at from here on
uint32_t dispatchResolved(const nrdk::PipelineKey& key, const void* constants, uint32_t constantsSize, const nrdcuTexture* textures, uint32_t texturesNum, uint32_t flags,
                          cudaStream_t s, uint32_t rowBegin, uint32_t rowEnd) {
    if (rowBegin % 16u) return fail(Result::INVALID_ARGUMENT, "nrdcuDispatchRows: rowBegin %u is not a multiple of 16", rowBegin);
    // a ragged rowEnd inside the frame would let the last CTA row ( 8 or 16 pixels tall ) store past it, into the neighbouring strip's rows
    if (rowEnd != 0xFFFFFFFFu && rowEnd % 16u) {
        uint32_t frameH = 0;
        for (uint32_t k = 0; k < texturesNum; k++) frameH = textures[k].height > frameH ? textures[k].height : frameH;
        if (rowEnd < frameH) return fail(Result::INVALID_ARGUMENT, "nrdcuDispatchRows: rowEnd %u is inside the frame and not a multiple of 16", rowEnd);
    }
    nrdk::Rows rows;
    rows.begin = (int)rowBegin;
    rows.end = rowEnd > 0x7FFFFFFFu ? 0x7FFFFFFF : (int)rowEnd;
    const bool partial = rowBegin != 0 || rowEnd != 0xFFFFFFFFu;
    std::string err;
    uint32_t r = 0;
    switch (key.family) {
        case nrdk::FAMILY_CLEAR: {  // clears always cover the whole texture (frame 0 only)
            if (texturesNum != 1 || !textures[0].data) return fail(Result::INVALID_ARGUMENT, "Clear: exactly one texture expected");
            const nrdcuTexture& t = textures[0];
            uint32_t bpp = bytesPerTexel(t.format);
            if (!bpp) return fail(Result::UNSUPPORTED, "Clear: unsupported format %u", t.format);
            nrdk::launchClear(t.data, (int)(t.width * bpp), (int)t.height, (int)t.pitchBytes, s);
            return checkLaunch("Clear") ? (uint32_t)Result::SUCCESS : (uint32_t)Result::FAILURE;
        }
        case nrdk::FAMILY_REBLUR: return dispatchReblur(key, constants, constantsSize, textures, texturesNum, flags, rows, s);
        case nrdk::FAMILY_SIGMA: r = nrdk::dispatchSigma(key, constants, constantsSize, textures, texturesNum, rows, s, err); break;
        case nrdk::FAMILY_RELAX: r = nrdk::dispatchRelax(key, constants, constantsSize, textures, texturesNum, rows, s, err); break;
        case nrdk::FAMILY_REFERENCE:
            if (partial) return fail(Result::UNSUPPORTED, "%s: the REFERENCE denoiser has no row ranges (multi-GPU strips)", key.id);
            r = nrdk::dispatchReference(key, constants, constantsSize, textures, texturesNum, 0u, 0u, s, err);
            break;
        default: return fail(Result::UNSUPPORTED, "no CUDA kernel for shader '%s'", key.id);
    }
    if (r != (uint32_t)Result::SUCCESS) return fail((Result)r, "%s", err.c_str());
    return checkLaunch(key.id) ? (uint32_t)Result::SUCCESS : (uint32_t)Result::FAILURE;
}

}  // namespace

namespace nrdk {
thread_local GraphReplay* g_graphReplay = nullptr;
thread_local GraphRecord* g_graphRecord = nullptr;
thread_local bool g_pdlAllowed = false;   // launchK may launch with programmatic stream serialization: whole frames of a context that is not a strip of a tiled frame
void countLaunch() { g_launchCount.fetch_add(1, std::memory_order_relaxed); }  // kernels launched by frontend.cu
}

// =================================================================================================================
// C ABI
// =================================================================================================================
extern "C" {

NRDCU_API const char* nrdcuGetLastError(void) { return g_lastError.c_str(); }
NRDCU_API uint64_t nrdcuGetLaunchCount(void) { return g_launchCount.load(); }
NRDCU_API uint32_t nrdcuGetMirrorProbe(uint64_t* out, int reset) {
    unsigned long long v[14] = {};
    if (!nrdk::readMirrorProbe(v, reset != 0)) return fail(Result::FAILURE, "nrdcuGetMirrorProbe: %s", cudaGetErrorString(cudaGetLastError()));
    if (out) for (int i = 0; i < 14; i++) out[i] = v[i];
    return 0;
}

NRDCU_API uint32_t nrdcuDispatchRows(const char* shaderIdentifier, const void* constants, uint32_t constantsSize, const nrdcuTexture* textures, uint32_t texturesNum,
                                     uint32_t flags, void* stream, uint32_t rowBegin, uint32_t rowEnd) {
    if (!shaderIdentifier || (!textures && texturesNum)) return fail(Result::INVALID_ARGUMENT, "nrdcuDispatch: null argument");
    // the context-less entry resolves the identifier on every call; a context ( nrdcuCreate ) does it once per pipeline
    return dispatchResolved(nrdk::resolvePipeline(shaderIdentifier), constants, constantsSize, textures, texturesNum, flags, (cudaStream_t)stream, rowBegin, rowEnd);
}

NRDCU_API uint32_t nrdcuDispatch(const char* shaderIdentifier, const void* constants, uint32_t constantsSize, const nrdcuTexture* textures, uint32_t texturesNum,
                                 uint32_t flags, void* stream) {
    return nrdcuDispatchRows(shaderIdentifier, constants, constantsSize, textures, texturesNum, flags, stream, 0u, 0xFFFFFFFFu);
}

}  // extern "C"

// =================================================================================================================
// Context: nrd::Instance + pools
// =================================================================================================================
struct HostBinding {
    void* host = nullptr;
    nrdcuTexture device = {};     // what the kernels bind (inputs: buffer 0)
    nrdcuTexture deviceAlt = {};  // pipelined path: input buffer 1 / output staging the D2H copy reads from
    uint32_t hostPitch = 0;
    int direction = 0;
    bool used = false;
};

struct nrdcuContext {
    Instance* instance = nullptr;
    int device = 0;
    uint32_t flags = 0;
    uint16_t width = 0, height = 0;
    std::vector<nrdcuTexture> permanent, transient;
    std::vector<uint32_t> permanentDs, transientDs;   // TextureDesc::downsampleFactor of each pool texture
    GeomPlane plane;                                  // REBLUR's per-frame geometry plane ( allocated when the instance holds a REBLUR denoiser )
    std::vector<nrdk::PipelineKey> pipelines;         // InstanceDesc::pipelines resolved to kernels at creation
    std::vector<DenoiserDesc> denoisers;              // ( identifier, denoiser ) pairs of the instance
    std::vector<uint8_t> checkerboard;                // per denoiser: its settings ask for half-width ( checkerboarded ) radiance inputs
    bool halfWidthInputs = false;
    nrdcuTexture user[(size_t)ResourceType::MAX_NUM] = {};
    HostBinding hostBindings[(size_t)ResourceType::MAX_NUM];
    std::vector<void*> allocations;
    uint64_t poolBytes = 0;
    std::vector<nrdcuTexture> scratch;
    std::vector<uint8_t> scratchIsStorage;
    // pipelined host path (nrdcuDenoiseHostPipelined): copies on their own streams, double-buffered inputs
    struct HostPipe {
        cudaStream_t h2d = nullptr, d2h = nullptr;
        cudaEvent_t inReady[2] = {}, inConsumed[2] = {}, outCopied = nullptr, d2hDone = nullptr;
        bool consumedValid[2] = {false, false}, d2hValid = false;
        uint64_t frame = 0;
    } pipe;
    // host frames ( nrdcuHostFrame* ): one contiguous device block per direction, same layout as the pinned host blocks
    struct FrameSlot { uint32_t resourceType; nrdcuTexture tex; size_t offset; };   // tex.data is unused here: views are base + offset
    struct FrameLayout { std::vector<FrameSlot> slots; size_t bytes = 0; } frameLayout[2];
    uint8_t* frameIn[2] = {nullptr, nullptr};    // double-buffered inputs
    uint8_t *frameOut = nullptr, *frameOutStage = nullptr;
    // multi-GPU strips over peer memory (nrdcuTile*): textures of the two neighbouring strips mapped through CUDA IPC
    struct HaloRule { std::string pass; uint32_t binding, rows; };
    struct Tile {
        bool attached = false;
        std::vector<nrdcuTexture> exported;   // permanent pool, transient pool, shared user textures — the order of the export blob
        std::vector<void*> peerBase[2];       // [0] = neighbour above, [1] = below: base of its copy of exported[i]
        uint32_t* flags = nullptr;            // this GPU's two flag words: [0] written by the neighbour above, [1] below
        uint32_t* peerFlags[2] = {nullptr, nullptr};
        uint32_t* hostError = nullptr;        // pinned, mapped: set by the wait kernel on timeout
        uint32_t seq = 0, defaultHalo = 64;
        cudaStream_t seamStream = nullptr;    // seam rows first: their push overlaps the interior of the same pass
        cudaEvent_t seamsDone = nullptr, pushed = nullptr;
        bool overlap = true;
        std::vector<HaloRule> rules;
        uint64_t bytesPushed = 0;
    } tile;
    // NRDCU_FLAG_CUDA_GRAPH: frames kept as instantiated graphs, keyed by what they bind ( ping-pong parity, frame 0 with its clears, ... )
    struct FrameGraph {
        uint64_t signature = 0, lastUse = 0;
        cudaGraph_t graph = nullptr;          // kept alive: the node handles below belong to it
        cudaGraphExec_t exec = nullptr;
        std::vector<cudaGraphNode_t> nodes;   // kernel nodes in launch order
        std::vector<const void*> funcs;       // ... and the kernel each one runs
    };
    std::vector<FrameGraph> graphs;
    cudaStream_t captureStream = nullptr;     // frames are captured on a stream of the library's own ( the caller's may be the legacy default stream )
    uint64_t graphClock = 0, graphCaptures = 0, graphReplays = 0;
    // strips: history is fetched at pixel + motion, the apron covers `defaultHalo - 2` rows of it. Checked on the device every frame ( kernels/peer_halo.cu )
    float motionScaleYRows = 0.0f;
    bool motionIsWorldSpace = false;
    uint32_t rectHeight = 0;
    uint32_t* motionExcess = nullptr;        // pinned, mapped: the worst number of rows a history fetch landed beyond the apron, over all frames so far
    // per-dispatch CUDA-event timing (bench.py's live roofline measurement)
    struct ProfileEntry { const char* name; double totalMs = 0; uint64_t count = 0; };
    struct PendingTiming { const char* name; cudaEvent_t start, stop; };
    bool profiling = false;
    std::vector<ProfileEntry> profile;
    std::vector<PendingTiming> pending;
    std::vector<cudaEvent_t> freeEvents;
};

namespace {

void ensurePipe(nrdcuContext* ctx);   // copy streams + events of the pipelined host paths ( defined with the host frames below )

bool allocTexture(nrdcuContext* ctx, uint32_t fmt, uint32_t w, uint32_t h, nrdcuTexture& out, bool countAsPool) {
    uint32_t bpp = bytesPerTexel(fmt);
    if (!bpp) {
        g_lastError = "unsupported pool format " + std::to_string(fmt);
        return false;
    }
    // 256-byte row pitch: every row starts on a full L2 sector group and satisfies TMA / 128-bit vector alignment
    uint32_t pitch = (w * bpp + 255u) & ~255u;
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, (size_t)pitch * h);
    if (e != cudaSuccess) {
        g_lastError = std::string("cudaMalloc: ") + cudaGetErrorString(e);
        return false;
    }
    cudaMemset(p, 0, (size_t)pitch * h);
    ctx->allocations.push_back(p);
    if (countAsPool) ctx->poolBytes += (uint64_t)pitch * h;
    out = {p, w, h, pitch, fmt};
    return true;
}

// Seam traffic of one dispatch in peer mode: rows of every written texture that the neighbours' later passes reach into go straight
// into the neighbours' copies, then flag + wait (kernels/peer_halo.cu). Every dispatch signals and waits, with or without rows to
// push: that keeps neighbouring strips within one pass of each other, which is what makes writing into their aprons safe.
// rows of apron the written textures of `dd` need on the neighbours ( the largest over its storage bindings; 0 for clears )
uint32_t maxHaloRows(const nrdcuContext* ctx, const DispatchDesc& dd) {
    const nrdcuContext::Tile& t = ctx->tile;
    if (!strncmp(dd.name, "Clear", 5)) return 0;
    const char* dash = strstr(dd.name, " - ");
    const char* passName = dash ? dash + 3 : dd.name;
    uint32_t most = 0;
    for (uint32_t k = 0; k < dd.resourcesNum; k++) {
        if (dd.resources[k].descriptorType != DescriptorType::STORAGE_TEXTURE) continue;
        uint32_t halo = t.defaultHalo;
        for (const auto& r : t.rules)
            if (r.binding == k && r.pass == passName) halo = r.rows < t.defaultHalo ? r.rows : t.defaultHalo;
        most = std::max(most, halo);
    }
    return most;
}

// `pushStream`: where the seam rows are stored into the neighbours and the flag is raised ( the launch stream, or the context's seam stream when the
// pass was split into seam rows first / interior second ); the wait for the neighbours' flags always goes onto the launch stream
uint32_t pushHalos(nrdcuContext* ctx, const DispatchDesc& dd, uint32_t rowBegin, uint32_t rowEnd, cudaStream_t stream, cudaStream_t pushStream) {
    nrdcuContext::Tile& t = ctx->tile;
    const uint32_t fullH = ctx->height;
    const uint32_t y0 = rowBegin, y1 = rowEnd < fullH ? rowEnd : fullH;
    const char* dash = strstr(dd.name, " - ");
    const char* passName = dash ? dash + 3 : dd.name;
    nrdk::HaloSegments segs;
    segs.n = 0;
    for (uint32_t k = 0; k < dd.resourcesNum; k++) {
        if (dd.resources[k].descriptorType != DescriptorType::STORAGE_TEXTURE) continue;
        uint32_t halo = t.defaultHalo;
        for (const auto& r : t.rules)
            if (r.binding == k && r.pass == passName) halo = r.rows < t.defaultHalo ? r.rows : t.defaultHalo;
        if (!halo || !strncmp(dd.name, "Clear", 5)) continue;  // clears cover the whole texture on every GPU
        const nrdcuTexture& tex = ctx->scratch[k];
        size_t idx = t.exported.size();
        for (size_t i = 0; i < t.exported.size(); i++)
            if (t.exported[i].data == tex.data) idx = i;
        if (idx == t.exported.size())
            return fail(Result::INVALID_ARGUMENT, "'%s' writes binding %u into a texture the neighbouring strips cannot see: in peer mode user outputs must come from nrdcuAllocSharedTexture",
                        dd.name, k);
        // the pool's own TextureDesc::downsampleFactor ( re-deriving it from the two heights is wrong for e.g. 130 rows -> 9 tile rows )
        uint32_t ds = 1;
        for (size_t i = 0; i < ctx->permanent.size(); i++) if (ctx->permanent[i].data == tex.data) ds = ctx->permanentDs[i];
        for (size_t i = 0; i < ctx->transient.size(); i++) if (ctx->transient[i].data == tex.data) ds = ctx->transientDs[i];
        const uint32_t ty0 = y0 / ds, ty1 = std::min<uint32_t>((y1 + ds - 1) / ds, tex.height), h = (halo + ds - 1) / ds;
        const uint32_t up1 = std::min(ty0 + h, ty1), down0 = ty1 > ty0 + h ? ty1 - h : ty0;
        const uint32_t range[2][2] = {{ty0, up1}, {down0, ty1}};  // rows for the neighbour above / below
        for (int d = 0; d < 2; d++) {
            if (!t.peerFlags[d] || range[d][1] <= range[d][0]) continue;
            if (segs.n >= nrdk::kMaxHaloSegments) return fail(Result::FAILURE, "'%s': too many halo segments", dd.name);
            const size_t off = (size_t)range[d][0] * tex.pitchBytes, bytes = (size_t)(range[d][1] - range[d][0]) * tex.pitchBytes;
            segs.s[segs.n++] = {(const uint8_t*)tex.data + off, (uint8_t*)t.peerBase[d][idx] + off, (uint32_t)(bytes / 16), 0u};
            t.bytesPushed += bytes;
        }
    }
    t.seq++;
    nrdk::launchHaloPush(segs, pushStream);
    nrdk::launchHaloSignal(t.peerFlags[0] ? t.peerFlags[0] + 1 : nullptr, t.peerFlags[1] ? t.peerFlags[1] + 0 : nullptr, t.seq, pushStream);
    if (pushStream != stream) {   // the next pass may overwrite what is being pushed ( TEMP1 / TEMP2 alternate ): it starts after the push has left
        cudaEventRecord(t.pushed, pushStream);
        cudaStreamWaitEvent(stream, t.pushed, 0);
    }
    nrdk::launchHaloWait(t.flags, t.seq, t.peerFlags[0] != nullptr, t.peerFlags[1] != nullptr, t.hostError, stream);
    if (!checkLaunch("halo push")) return (uint32_t)Result::FAILURE;
    return 0;
}

struct TileExportEntry {
    cudaIpcMemHandle_t handle;
    uint64_t bytes;
};

}  // namespace

extern "C" {

// ---- strips over peer memory ---------------------------------------------------------------------------------------
NRDCU_API uint32_t nrdcuAllocSharedTexture(nrdcuContext* ctx, uint32_t format, uint32_t width, uint32_t height, nrdcuTexture* out) {
    if (!ctx || !out) return fail(Result::INVALID_ARGUMENT, "nrdcuAllocSharedTexture: null argument");
    if (ctx->tile.attached) return fail(Result::INVALID_ARGUMENT, "nrdcuAllocSharedTexture: allocate before nrdcuTileExport / nrdcuTileAttach");
    cudaSetDevice(ctx->device);
    nrdcuTexture t = {};
    if (!allocTexture(ctx, format, width, height, t, false)) return (uint32_t)Result::FAILURE;
    ctx->tile.exported.push_back(t);  // appended behind the pools when the export list is built
    *out = t;
    return 0;
}

static void buildExportList(nrdcuContext* ctx) {
    std::vector<nrdcuTexture> shared;
    shared.swap(ctx->tile.exported);
    // keep only genuine shared textures (anything that is not already a pool texture), pools first
    std::vector<nrdcuTexture> list = ctx->permanent;
    list.insert(list.end(), ctx->transient.begin(), ctx->transient.end());
    const size_t pools = list.size();
    for (const nrdcuTexture& s : shared) {
        bool isPool = false;
        for (size_t i = 0; i < pools; i++) isPool |= list[i].data == s.data;
        if (!isPool) list.push_back(s);
    }
    ctx->tile.exported.swap(list);
}

NRDCU_API uint32_t nrdcuTileExportSize(nrdcuContext* ctx) {
    if (!ctx) return 0;
    buildExportList(ctx);
    return (uint32_t)(sizeof(uint32_t) * 2 + (ctx->tile.exported.size() + 1) * sizeof(TileExportEntry));
}

NRDCU_API uint32_t nrdcuTileExport(nrdcuContext* ctx, void* blob, uint32_t blobSize) {
    if (!ctx || !blob) return fail(Result::INVALID_ARGUMENT, "nrdcuTileExport: null argument");
    if (blobSize < nrdcuTileExportSize(ctx)) return fail(Result::INVALID_ARGUMENT, "nrdcuTileExport: blob too small");
    cudaSetDevice(ctx->device);
    nrdcuContext::Tile& t = ctx->tile;
    if (!t.flags) {
        void* p = nullptr;
        if (cudaMalloc(&p, 256) != cudaSuccess) return fail(Result::FAILURE, "nrdcuTileExport: cudaMalloc failed");
        cudaMemset(p, 0, 256);
        ctx->allocations.push_back(p);
        t.flags = (uint32_t*)p;
        if (cudaHostAlloc((void**)&t.hostError, sizeof(uint32_t), cudaHostAllocMapped) != cudaSuccess) return fail(Result::FAILURE, "nrdcuTileExport: cudaHostAlloc failed");
        *t.hostError = 0;
    }
    uint8_t* w = (uint8_t*)blob;
    const uint32_t header[2] = {0x4E524454u, (uint32_t)t.exported.size()};
    memcpy(w, header, sizeof(header));
    w += sizeof(header);
    for (size_t i = 0; i <= t.exported.size(); i++) {
        TileExportEntry e = {};
        void* ptr = i < t.exported.size() ? t.exported[i].data : (void*)t.flags;
        e.bytes = i < t.exported.size() ? (uint64_t)t.exported[i].pitchBytes * t.exported[i].height : 256;
        cudaError_t err = cudaIpcGetMemHandle(&e.handle, ptr);
        if (err != cudaSuccess) return fail(Result::FAILURE, "cudaIpcGetMemHandle: %s", cudaGetErrorString(err));
        memcpy(w, &e, sizeof(e));
        w += sizeof(e);
    }
    cudaDeviceSynchronize();  // the memset of the flag words has landed before anybody maps them
    return 0;
}

// blobAbove / blobBelow: the export blobs of the ranks owning the strips above / below this one (NULL at the frame edges)
NRDCU_API uint32_t nrdcuTileAttach(nrdcuContext* ctx, const void* blobAbove, const void* blobBelow, uint32_t blobSize) {
    if (!ctx) return fail(Result::INVALID_ARGUMENT, "nrdcuTileAttach: null context");
    nrdcuContext::Tile& t = ctx->tile;
    if (!t.flags) return fail(Result::INVALID_ARGUMENT, "nrdcuTileAttach: call nrdcuTileExport first");
    cudaSetDevice(ctx->device);
    const void* blobs[2] = {blobAbove, blobBelow};
    for (int d = 0; d < 2; d++) {
        if (!blobs[d]) continue;
        const uint8_t* r = (const uint8_t*)blobs[d];
        uint32_t header[2];
        memcpy(header, r, sizeof(header));
        r += sizeof(header);
        if (header[0] != 0x4E524454u || header[1] != t.exported.size() || blobSize < sizeof(header) + (header[1] + 1) * sizeof(TileExportEntry))
            return fail(Result::INVALID_ARGUMENT, "nrdcuTileAttach: neighbour blob does not match this context (%u textures vs %zu)", header[1], t.exported.size());
        t.peerBase[d].assign(t.exported.size(), nullptr);
        for (size_t i = 0; i <= t.exported.size(); i++) {
            TileExportEntry e;
            memcpy(&e, r, sizeof(e));
            r += sizeof(e);
            void* mapped = nullptr;
            cudaError_t err = cudaIpcOpenMemHandle(&mapped, e.handle, cudaIpcMemLazyEnablePeerAccess);
            if (err != cudaSuccess) return fail(Result::FAILURE, "cudaIpcOpenMemHandle (texture %zu of the neighbour %s): %s", i, d ? "below" : "above", cudaGetErrorString(err));
            if (i < t.exported.size())
                t.peerBase[d][i] = mapped;
            else
                t.peerFlags[d] = (uint32_t*)mapped;
        }
    }
    if (!t.seamStream) {
        // highest priority: the block scheduler hands freed SM slots to the push kernel first, so the seam rows leave while the interior still runs
        int leastPriority = 0, greatestPriority = 0;
        cudaDeviceGetStreamPriorityRange(&leastPriority, &greatestPriority);
        cudaStreamCreateWithPriority(&t.seamStream, cudaStreamNonBlocking, greatestPriority);
        cudaEventCreateWithFlags(&t.seamsDone, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&t.pushed, cudaEventDisableTiming);
    }
    t.overlap = !(ctx->flags & NRDCU_FLAG_NO_SEAM_OVERLAP);
    t.attached = true;
    return 0;
}

NRDCU_API uint32_t nrdcuTileSetHalo(nrdcuContext* ctx, const char* passName, uint32_t binding, uint32_t rows) {
    if (!ctx) return fail(Result::INVALID_ARGUMENT, "nrdcuTileSetHalo: null context");
    if (!passName)
        ctx->tile.defaultHalo = rows;
    else
        ctx->tile.rules.push_back({passName, binding, rows});
    return 0;
}

NRDCU_API uint32_t nrdcuTileGetStatus(nrdcuContext* ctx, uint64_t* bytesPushed, uint32_t* error) {
    if (!ctx) return fail(Result::INVALID_ARGUMENT, "nrdcuTileGetStatus: null context");
    if (bytesPushed) *bytesPushed = ctx->tile.bytesPushed;
    if (error) *error = ctx->tile.hostError ? *(volatile uint32_t*)ctx->tile.hostError : 0u;
    return 0;
}

NRDCU_API uint32_t nrdcuTileGetMotionBound(nrdcuContext* ctx, uint32_t* boundRows, uint32_t* worstExcessRows) {
    if (!ctx) return fail(Result::INVALID_ARGUMENT, "nrdcuTileGetMotionBound: null context");
    if (boundRows) *boundRows = ctx->tile.defaultHalo >= 2 ? ctx->tile.defaultHalo - 2 : 0u;
    if (worstExcessRows) *worstExcessRows = ctx->motionExcess ? *(volatile uint32_t*)ctx->motionExcess : 0u;
    return 0;
}

NRDCU_API uint32_t nrdcuCreate(const void* instanceCreationDesc, uint16_t resourceWidth, uint16_t resourceHeight, int device, uint32_t flags, nrdcuContext** out) {
    if (!instanceCreationDesc || !out || !resourceWidth || !resourceHeight) return fail(Result::INVALID_ARGUMENT, "nrdcuCreate: null or zero argument");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) return fail(Result::FAILURE, "nrdcuCreate: no CUDA device (%s) — this library has no CPU fallback", cudaGetErrorString(e));
    if (device < 0 || device >= count) return fail(Result::INVALID_ARGUMENT, "nrdcuCreate: device %d out of range (%d devices)", device, count);
    if ((e = cudaSetDevice(device)) != cudaSuccess) return fail(Result::FAILURE, "cudaSetDevice: %s", cudaGetErrorString(e));

    nrdcuContext* ctx = new nrdcuContext();
    ctx->device = device;
    ctx->flags = flags;
    ctx->width = resourceWidth;
    ctx->height = resourceHeight;
    Result r = CreateInstance(*(const InstanceCreationDesc*)instanceCreationDesc, ctx->instance);
    if (r != Result::SUCCESS) {
        delete ctx;
        return fail(r, "nrd::CreateInstance failed (%u)", (uint32_t)r);
    }
    const InstanceCreationDesc& icd = *(const InstanceCreationDesc*)instanceCreationDesc;
    ctx->denoisers.assign(icd.denoisers, icd.denoisers + icd.denoisersNum);
    ctx->checkerboard.assign(icd.denoisersNum, 0);
    const InstanceDesc& d = *GetInstanceDesc(*ctx->instance);
    // pipelines are picked HERE, once ( the reference integration creates its pipeline objects in Recreate, NRDIntegration.hpp:199-239 ): every
    // shaderIdentifier of the instance becomes a PipelineKey, and an identifier without a kernel fails the creation instead of the first frame
    ctx->pipelines.resize(d.pipelinesNum);
    for (uint32_t i = 0; i < d.pipelinesNum; i++) {
        ctx->pipelines[i] = nrdk::resolvePipeline(d.pipelines[i].shaderIdentifier);
        if (ctx->pipelines[i].family == nrdk::FAMILY_UNKNOWN) {
            const std::string id = d.pipelines[i].shaderIdentifier;
            nrdcuDestroy(ctx);
            return fail(Result::UNSUPPORTED, "nrdcuCreate: no CUDA kernel for pipeline %u '%s'", i, id.c_str());
        }
    }
    auto makePool = [&](const TextureDesc* descs, uint32_t n, std::vector<nrdcuTexture>& pool, std::vector<uint32_t>& dsOut) {
        pool.resize(n);
        dsOut.resize(n);
        for (uint32_t i = 0; i < n; i++) {
            uint32_t ds = descs[i].downsampleFactor;
            dsOut[i] = ds;
            if (!allocTexture(ctx, (uint32_t)descs[i].format, (resourceWidth + ds - 1) / ds, (resourceHeight + ds - 1) / ds, pool[i], true)) return false;
        }
        return true;
    };
    if (!makePool(d.permanentPool, d.permanentPoolSize, ctx->permanent, ctx->permanentDs) || !makePool(d.transientPool, d.transientPoolSize, ctx->transient, ctx->transientDs)) {
        nrdcuDestroy(ctx);
        return (uint32_t)Result::FAILURE;
    }
    for (const DenoiserDesc& dn : ctx->denoisers)
        if (dn.denoiser <= Denoiser::REBLUR_DIFFUSE_DIRECTIONAL_OCCLUSION && !ctx->plane.tex.data &&
            !allocTexture(ctx, (uint32_t)Format::RGBA32_SFLOAT, resourceWidth, resourceHeight, ctx->plane.tex, true)) {
            nrdcuDestroy(ctx);
            return (uint32_t)Result::FAILURE;
        }
    *out = ctx;
    return (uint32_t)Result::SUCCESS;
}

NRDCU_API void nrdcuDestroy(nrdcuContext* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (int d = 0; d < 2; d++) {
        for (void* p : ctx->tile.peerBase[d])
            if (p) cudaIpcCloseMemHandle(p);
        if (ctx->tile.peerFlags[d]) cudaIpcCloseMemHandle(ctx->tile.peerFlags[d]);
    }
    for (nrdcuContext::FrameGraph& fg : ctx->graphs) {
        cudaGraphExecDestroy(fg.exec);
        cudaGraphDestroy(fg.graph);
    }
    if (ctx->captureStream) cudaStreamDestroy(ctx->captureStream);
    if (ctx->tile.hostError) cudaFreeHost(ctx->tile.hostError);
    if (ctx->motionExcess) cudaFreeHost(ctx->motionExcess);
    if (ctx->tile.seamStream) {
        cudaStreamDestroy(ctx->tile.seamStream);
        cudaEventDestroy(ctx->tile.seamsDone);
        cudaEventDestroy(ctx->tile.pushed);
    }
    if (ctx->pipe.h2d) {
        cudaStreamDestroy(ctx->pipe.h2d);
        cudaStreamDestroy(ctx->pipe.d2h);
        for (int i = 0; i < 2; i++) {
            cudaEventDestroy(ctx->pipe.inReady[i]);
            cudaEventDestroy(ctx->pipe.inConsumed[i]);
        }
        cudaEventDestroy(ctx->pipe.outCopied);
        cudaEventDestroy(ctx->pipe.d2hDone);
    }
    for (void* p : ctx->allocations) cudaFree(p);
    if (ctx->instance) DestroyInstance(*ctx->instance);
    delete ctx;
}

NRDCU_API void* nrdcuGetInstance(nrdcuContext* ctx) { return ctx ? ctx->instance : nullptr; }
NRDCU_API uint32_t nrdcuGetMemoryUsage(nrdcuContext* ctx, uint64_t* persistentBytes, uint64_t* aliasableBytes, uint64_t* privateBytes) {
    if (!ctx) return fail(Result::INVALID_ARGUMENT, "nrdcuGetMemoryUsage: null context");
    uint64_t perm = 0, tran = 0;
    for (const nrdcuTexture& t : ctx->permanent) perm += (uint64_t)t.pitchBytes * t.height;
    for (const nrdcuTexture& t : ctx->transient) tran += (uint64_t)t.pitchBytes * t.height;
    if (persistentBytes) *persistentBytes = perm;
    if (aliasableBytes) *aliasableBytes = tran;
    if (privateBytes) *privateBytes = ctx->poolBytes - perm - tran;   // the geometry plane: not part of the reference's pools
    return 0;
}
NRDCU_API uint64_t nrdcuGetPoolBytes(nrdcuContext* ctx) { return ctx ? ctx->poolBytes : 0; }

NRDCU_API uint32_t nrdcuSetCommonSettings(nrdcuContext* ctx, const void* commonSettings) {
    if (!ctx || !commonSettings) return fail(Result::INVALID_ARGUMENT, "nrdcuSetCommonSettings: null argument");
    const CommonSettings& cs = *(const CommonSettings*)commonSettings;
    if (cs.resourceSize[0] != ctx->width || cs.resourceSize[1] != ctx->height)
        return fail(Result::INVALID_ARGUMENT, "resourceSize %ux%u does not match the pools created at %ux%u", cs.resourceSize[0], cs.resourceSize[1], ctx->width, ctx->height);
    Result r = SetCommonSettings(*ctx->instance, cs);
    if (r == Result::SUCCESS) {   // what the strips' motion bound check needs ( nrdcuDenoiseRows )
        ctx->motionScaleYRows = cs.motionVectorScale[1] * (float)cs.rectSize[1];
        ctx->motionIsWorldSpace = cs.isMotionVectorInWorldSpace;
        ctx->rectHeight = cs.rectSize[1];
    }
    return r == Result::SUCCESS ? 0u : fail(r, "nrd::SetCommonSettings rejected the settings");
}

NRDCU_API uint32_t nrdcuSetDenoiserSettings(nrdcuContext* ctx, uint32_t identifier, const void* denoiserSettings) {
    if (!ctx || !denoiserSettings) return fail(Result::INVALID_ARGUMENT, "nrdcuSetDenoiserSettings: null argument");
    Result r = SetDenoiserSettings(*ctx->instance, identifier, denoiserSettings);
    if (r != Result::SUCCESS) return fail(r, "nrd::SetDenoiserSettings: unknown identifier %u", identifier);
    // remember whether the radiance inputs are checkerboarded ( half width ): decides how small a user texture may be ( nrdcuDenoiseRows )
    ctx->halfWidthInputs = false;
    for (size_t i = 0; i < ctx->denoisers.size(); i++) {
        const Denoiser dn = ctx->denoisers[i].denoiser;
        if (ctx->denoisers[i].identifier == identifier) {
            if (dn <= Denoiser::REBLUR_DIFFUSE_DIRECTIONAL_OCCLUSION) ctx->checkerboard[i] = ((const ReblurSettings*)denoiserSettings)->checkerboardMode != CheckerboardMode::OFF;
            else if (dn <= Denoiser::RELAX_DIFFUSE_SPECULAR_SH) ctx->checkerboard[i] = ((const RelaxSettings*)denoiserSettings)->checkerboardMode != CheckerboardMode::OFF;
        }
        ctx->halfWidthInputs |= ctx->checkerboard[i] != 0;
    }
    return 0u;
}

NRDCU_API uint32_t nrdcuSetResource(nrdcuContext* ctx, uint32_t resourceType, const nrdcuTexture* texture) {
    if (!ctx || !texture || resourceType >= (uint32_t)ResourceType::TRANSIENT_POOL) return fail(Result::INVALID_ARGUMENT, "nrdcuSetResource: bad slot %u", resourceType);
    if (!texture->data || !bytesPerTexel(texture->format)) return fail(Result::INVALID_ARGUMENT, "nrdcuSetResource: null data or unsupported format %u", texture->format);
    ctx->user[resourceType] = *texture;
    return 0;
}

NRDCU_API uint32_t nrdcuGetPoolTexture(nrdcuContext* ctx, int isPermanent, uint32_t index, nrdcuTexture* out) {
    if (!ctx || !out) return fail(Result::INVALID_ARGUMENT, "nrdcuGetPoolTexture: null argument");
    const std::vector<nrdcuTexture>& pool = isPermanent ? ctx->permanent : ctx->transient;
    if (index >= pool.size()) return fail(Result::INVALID_ARGUMENT, "pool index %u out of range", index);
    *out = pool[index];
    return 0;
}

NRDCU_API uint32_t nrdcuDenoise(nrdcuContext* ctx, const uint32_t* identifiers, uint32_t identifiersNum, void* stream) {
    return nrdcuDenoiseRows(ctx, identifiers, identifiersNum, stream, 0u, 0xFFFFFFFFu, nullptr, nullptr);
}

NRDCU_API uint32_t nrdcuDenoiseRows(nrdcuContext* ctx, const uint32_t* identifiers, uint32_t identifiersNum, void* stream, uint32_t rowBegin, uint32_t rowEnd,
                                    nrdcuDispatchCallback afterDispatch, void* userArg) {
    if (!ctx) return fail(Result::INVALID_ARGUMENT, "nrdcuDenoise: null context");
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return fail(Result::FAILURE, "cudaSetDevice: %s", cudaGetErrorString(e));
    const DispatchDesc* dispatches = nullptr;
    uint32_t n = 0;
    Result r = GetComputeDispatches(*ctx->instance, identifiers, identifiersNum, dispatches, n);
    if (r != Result::SUCCESS) return fail(r, "nrd::GetComputeDispatches failed (%u)", (uint32_t)r);
    // new frame, new G-buffer: the geometry plane is decoded again by the first spatial pass that needs it
    ctx->plane.row0 = ctx->plane.row1 = 0;
    ctx->plane.viewZCopy = nullptr;
    ctx->plane.margin = ctx->tile.defaultHalo;
    ctx->plane.stripRow0 = (int)rowBegin;
    ctx->plane.stripRow1 = rowEnd > 0x7FFFFFFFu ? 0x7FFFFFFF : (int)rowEnd;
    struct PlaneScope {
        PlaneScope(GeomPlane* p) { g_plane = p; }
        ~PlaneScope() { g_plane = nullptr; }
    } planeScope(&ctx->plane);
    // programmatic dependent launches for the frame of a stand-alone context; strips ( seam kernels on a second stream, flag / wait kernels of the peer exchange
    // between the passes ) and the context-less nrdcuDispatch keep plain stream serialization
    struct PdlScope {
        PdlScope(bool on) { nrdk::g_pdlAllowed = on; }
        ~PdlScope() { nrdk::g_pdlAllowed = false; }
    } pdlScope(!ctx->tile.attached);
    // SIGMA's Copy pass rides in its first blur pass while a context runs the frame ( kernels/sigma.cu ); per-dispatch callers and strips get every pass on its own
    struct SigmaFusionScope {
        SigmaFusionScope(bool on) { nrdk::sigmaSetCopyFusion(on); }
        ~SigmaFusionScope() { nrdk::sigmaSetCopyFusion(false); }
    } sigmaFusionScope(!afterDispatch && !ctx->tile.attached && rowBegin == 0 && rowEnd == 0xFFFFFFFFu);
    // a strip: does any history fetch of this frame leave the apron? ( 2D / 2.5D motion; world-space motion would need the reprojection itself )
    if ((rowBegin != 0 || rowEnd < ctx->height) && !ctx->motionIsWorldSpace && ctx->user[(uint32_t)ResourceType::IN_MV].data && ctx->tile.defaultHalo >= 2) {
        if (!ctx->motionExcess) {
            if (cudaHostAlloc((void**)&ctx->motionExcess, sizeof(uint32_t), cudaHostAllocMapped) != cudaSuccess) return fail(Result::FAILURE, "nrdcuDenoiseRows: cudaHostAlloc failed");
            *ctx->motionExcess = 0;
        }
        const nrdcuTexture& mv = ctx->user[(uint32_t)ResourceType::IN_MV];
        const int frameH = (int)(ctx->rectHeight ? ctx->rectHeight : ctx->height);
        nrdk::launchMotionBoundCheck(mv.data, mv.pitchBytes, mv.format, (int)std::min<uint32_t>(mv.width, ctx->width), (int)rowBegin, (int)std::min<uint32_t>(rowEnd, (uint32_t)frameH), frameH,
                                     ctx->motionScaleYRows, (int)ctx->tile.defaultHalo - 2, ctx->motionExcess, (cudaStream_t)stream);
    }
    // the frame: every dispatch of the list, in order, onto `stream`
    auto runFrame = [&](void* stream) -> uint32_t {
    for (uint32_t i = 0; i < n; i++) {
        const DispatchDesc& dd = dispatches[i];
        ctx->scratch.resize(dd.resourcesNum);
        for (uint32_t k = 0; k < dd.resourcesNum; k++) {
            const ResourceDesc& res = dd.resources[k];
            if (res.type == ResourceType::PERMANENT_POOL)
                ctx->scratch[k] = ctx->permanent[res.indexInPool];
            else if (res.type == ResourceType::TRANSIENT_POOL)
                ctx->scratch[k] = ctx->transient[res.indexInPool];
            else {
                ctx->scratch[k] = ctx->user[(uint32_t)res.type];
                if (!ctx->scratch[k].data) return fail(Result::INVALID_ARGUMENT, "'%s' needs user resource %s, which was not set", dd.name, GetResourceTypeString(res.type));
                // kernels fetch user inputs at coordinates clamped to the RECT, not to the texture: anything smaller than the resource size
                // would be read out of bounds ( guides — confidence / threshold mix — may have any size: they are sampled by uv )
                const bool guide = res.type == ResourceType::IN_DIFF_CONFIDENCE || res.type == ResourceType::IN_SPEC_CONFIDENCE || res.type == ResourceType::IN_DISOCCLUSION_THRESHOLD_MIX;
                if (!guide && (ctx->scratch[k].width * 2u < ctx->width || ctx->scratch[k].height < ctx->height ||
                               (ctx->scratch[k].width < ctx->width && !ctx->halfWidthInputs)))
                    return fail(Result::INVALID_ARGUMENT, "'%s': user resource %s is %ux%u, smaller than the %ux%u the instance was created for", dd.name, GetResourceTypeString(res.type),
                                ctx->scratch[k].width, ctx->scratch[k].height, ctx->width, ctx->height);
            }
        }
        cudaEvent_t evStart = nullptr, evStop = nullptr;
        if (ctx->profiling) {
            auto grab = [&]() {
                cudaEvent_t ev;
                if (!ctx->freeEvents.empty()) { ev = ctx->freeEvents.back(); ctx->freeEvents.pop_back(); }
                else cudaEventCreate(&ev);
                return ev;
            };
            evStart = grab();
            evStop = grab();
            cudaEventRecord(evStart, (cudaStream_t)stream);
        }
        const nrdk::PipelineKey& pipeline = ctx->pipelines[dd.pipelineIndex];
        auto run = [&](uint32_t r0, uint32_t r1) {
            return r1 > r0 ? dispatchResolved(pipeline, dd.constantBufferData, dd.constantBufferDataSize, ctx->scratch.data(), dd.resourcesNum, ctx->flags, (cudaStream_t)stream, r0, r1) : 0u;
        };
        // Strips over peer memory: the rows the neighbours need are computed FIRST, their push ( own stream ) then overlaps the interior of the same pass
        const uint32_t stripEnd = std::min<uint32_t>(rowEnd, ctx->height);
        uint32_t edge = ctx->tile.attached && ctx->tile.overlap ? (maxHaloRows(ctx, dd) + 15u) & ~15u : 0u;
        // three launches instead of one only pay when the interior is most of the strip: measured at 4K, strips of 540+ rows ( 2 / 4 GPUs ) are neutral, strips of
        // ~200 rows ( 8 GPUs ) lose 9 % to the extra launches ( 0.685 vs 0.625 ms per frame ) — there the pass is launched whole and the push follows it
        if (edge && stripEnd - rowBegin < 6u * edge) edge = 0;
        uint32_t rc = 0;
        if (edge) {
            const uint32_t topEnd = ctx->tile.peerFlags[0] ? rowBegin + edge : rowBegin;
            const uint32_t bottomBegin = ctx->tile.peerFlags[1] ? ((stripEnd - edge) & ~15u) : stripEnd;
            rc = run(rowBegin, topEnd);
            if (!rc) rc = run(bottomBegin, rowEnd);
            if (!rc) {
                cudaEventRecord(ctx->tile.seamsDone, (cudaStream_t)stream);
                cudaStreamWaitEvent(ctx->tile.seamStream, ctx->tile.seamsDone, 0);
                // ( the push + flag go onto the seam stream in pushHalos below; it waits for nothing but the seam rows )
            }
            if (!rc) rc = run(topEnd, bottomBegin);
        } else
            rc = run(rowBegin, rowEnd);
        if (ctx->profiling) {
            cudaEventRecord(evStop, (cudaStream_t)stream);
            ctx->pending.push_back({dd.name, evStart, evStop});   // kept on the error path too: nrdcuResolveProfile recycles the events
        }
        if (rc != 0) return rc;
        if (ctx->tile.attached) {
            rc = pushHalos(ctx, dd, rowBegin, rowEnd, (cudaStream_t)stream, edge ? ctx->tile.seamStream : (cudaStream_t)stream);
            if (rc != 0) return rc;
        }
        if (afterDispatch) {
            // which bindings the dispatch wrote (storage textures): what a strip has to trade with its neighbours
            ctx->scratchIsStorage.resize(dd.resourcesNum);
            for (uint32_t k = 0; k < dd.resourcesNum; k++) ctx->scratchIsStorage[k] = dd.resources[k].descriptorType == DescriptorType::STORAGE_TEXTURE ? 1 : 0;
            afterDispatch(userArg, i, dd.name, ctx->scratch.data(), ctx->scratchIsStorage.data(), dd.resourcesNum);
        }
    }
    nrdk::sigmaFlushPendingCopy((cudaStream_t)stream);   // ( a Copy no blur pass picked up: cannot happen with the reference's pass lists, harmless if it does )
    return 0;
    };
    auto resetPlane = [&]() {
        ctx->plane.row0 = ctx->plane.row1 = 0;
        ctx->plane.viewZCopy = nullptr;
    };
    const bool graphMode = (ctx->flags & NRDCU_FLAG_CUDA_GRAPH) && n != 0 && !ctx->tile.attached && !ctx->profiling && !afterDispatch && rowBegin == 0 && rowEnd == 0xFFFFFFFFu;
    if (!graphMode) return runFrame(stream);

    // ---- NRDCU_FLAG_CUDA_GRAPH ------------------------------------------------------------------------------------------------------------------
    // What decides the chain of kernels and everything in their parameters that is not a constant: which pipelines run and which memory they bind
    uint64_t signature = 1469598103934665603ull;
    auto mix = [&](uint64_t v) { signature = (signature ^ v) * 1099511628211ull; };
    mix(n);
    for (uint32_t i = 0; i < n; i++) {
        const DispatchDesc& dd = dispatches[i];
        mix(dd.pipelineIndex);
        mix(dd.resourcesNum);
        for (uint32_t k = 0; k < dd.resourcesNum; k++) {
            const ResourceDesc& res = dd.resources[k];
            const nrdcuTexture& t = res.type == ResourceType::PERMANENT_POOL ? ctx->permanent[res.indexInPool] : (res.type == ResourceType::TRANSIENT_POOL ? ctx->transient[res.indexInPool] : ctx->user[(uint32_t)res.type]);
            mix((uint64_t)(uintptr_t)t.data);
            mix(((uint64_t)t.width << 32) | t.height);
            mix(((uint64_t)t.pitchBytes << 32) | t.format);
        }
    }
    ctx->graphClock++;
    for (size_t g = 0; g < ctx->graphs.size(); g++) {
        nrdcuContext::FrameGraph& fg = ctx->graphs[g];
        if (fg.signature != signature) continue;
        // same chain as a frame seen before: run the host side of every dispatch with the launches redirected into the graph's kernel nodes
        nrdk::GraphReplay replay;
        replay.exec = fg.exec;
        replay.nodes = fg.nodes.data();
        replay.funcs = fg.funcs.data();
        replay.count = (uint32_t)fg.nodes.size();
        nrdk::g_graphReplay = &replay;
        uint32_t rc = runFrame(stream);
        nrdk::g_graphReplay = nullptr;
        if (rc != 0) return rc;
        if (!replay.failed && replay.cursor == replay.count) {
            cudaError_t ge = cudaGraphLaunch(fg.exec, (cudaStream_t)stream);
            if (ge != cudaSuccess) return fail(Result::FAILURE, "cudaGraphLaunch: %s", cudaGetErrorString(ge));
            fg.lastUse = ctx->graphClock;
            ctx->graphReplays++;
            return 0;
        }
        // the settings changed which kernels run ( another template instance, another number of launches ): forget this graph, capture again
        cudaGraphExecDestroy(fg.exec);
        cudaGraphDestroy(fg.graph);
        ctx->graphs.erase(ctx->graphs.begin() + (long)g);
        resetPlane();
        break;
    }
    if (!ctx->captureStream && (e = cudaStreamCreateWithFlags(&ctx->captureStream, cudaStreamNonBlocking)) != cudaSuccess) return fail(Result::FAILURE, "cudaStreamCreate: %s", cudaGetErrorString(e));
    nrdk::GraphRecord record;
    if ((e = cudaStreamBeginCapture(ctx->captureStream, cudaStreamCaptureModeThreadLocal)) != cudaSuccess) return fail(Result::FAILURE, "cudaStreamBeginCapture: %s", cudaGetErrorString(e));
    nrdk::g_graphRecord = &record;
    uint32_t rc = runFrame(ctx->captureStream);
    nrdk::g_graphRecord = nullptr;
    cudaGraph_t graph = nullptr;
    e = cudaStreamEndCapture(ctx->captureStream, &graph);
    if (rc != 0) {
        if (graph) cudaGraphDestroy(graph);
        return rc;
    }
    if (e != cudaSuccess || !graph) return fail(Result::FAILURE, "cudaStreamEndCapture: %s", cudaGetErrorString(e));
    nrdcuContext::FrameGraph fg;
    fg.signature = signature;
    fg.lastUse = ctx->graphClock;
    // the captured graph is one chain of kernel nodes: walk it from its root
    bool chain = !record.overflow;
    {
        size_t roots = 0, total = 0;
        cudaGraphGetNodes(graph, nullptr, &total);
        cudaGraphGetRootNodes(graph, nullptr, &roots);
        cudaGraphNode_t node = nullptr;
        if (roots == 1) {
            size_t one = 1;
            cudaGraphGetRootNodes(graph, &node, &one);
        } else
            chain = false;
        while (chain && node) {
            cudaGraphNodeType type;
            if (cudaGraphNodeGetType(node, &type) != cudaSuccess || type != cudaGraphNodeTypeKernel) chain = false;
            fg.nodes.push_back(node);
            size_t next = 0;
            cudaGraphNodeGetDependentNodes(node, nullptr, &next);
            if (next > 1) chain = false;
            cudaGraphNode_t nextNode = nullptr;
            if (next == 1) {
                size_t one = 1;
                cudaGraphNodeGetDependentNodes(node, &nextNode, &one);
            }
            node = nextNode;
        }
        if (fg.nodes.size() != total || fg.nodes.size() != record.count) chain = false;
    }
    e = cudaGraphInstantiate(&fg.exec, graph, 0);
    if (e != cudaSuccess) {
        cudaGraphDestroy(graph);
        return fail(Result::FAILURE, "cudaGraphInstantiate: %s", cudaGetErrorString(e));
    }
    fg.graph = graph;
    e = cudaGraphLaunch(fg.exec, (cudaStream_t)stream);
    if (e != cudaSuccess) {
        cudaGraphExecDestroy(fg.exec);
        cudaGraphDestroy(graph);
        return fail(Result::FAILURE, "cudaGraphLaunch: %s", cudaGetErrorString(e));
    }
    ctx->graphCaptures++;
    if (!chain) {   // not a plain chain of kernels: run it this once, keep nothing
        cudaStreamSynchronize((cudaStream_t)stream);
        cudaGraphExecDestroy(fg.exec);
        cudaGraphDestroy(graph);
        return 0;
    }
    fg.funcs.assign(record.funcs, record.funcs + record.count);
    if (ctx->graphs.size() >= 16) {   // least recently used out ( a renderer cycling through a few G-buffer sets times two ping-pong parities fits )
        size_t victim = 0;
        for (size_t g = 1; g < ctx->graphs.size(); g++)
            if (ctx->graphs[g].lastUse < ctx->graphs[victim].lastUse) victim = g;
        cudaStreamSynchronize((cudaStream_t)stream);
        cudaGraphExecDestroy(ctx->graphs[victim].exec);
        cudaGraphDestroy(ctx->graphs[victim].graph);
        ctx->graphs.erase(ctx->graphs.begin() + (long)victim);
    }
    ctx->graphs.push_back(std::move(fg));
    return 0;
}

NRDCU_API uint32_t nrdcuGetGraphStats(nrdcuContext* ctx, uint64_t* captures, uint64_t* replays, uint32_t* cached) {
    if (!ctx) return fail(Result::INVALID_ARGUMENT, "nrdcuGetGraphStats: null context");
    if (captures) *captures = ctx->graphCaptures;
    if (replays) *replays = ctx->graphReplays;
    if (cached) *cached = (uint32_t)ctx->graphs.size();
    return 0;
}


NRDCU_API uint32_t nrdcuSetProfiling(nrdcuContext* ctx, int enabled) {
    if (!ctx) return fail(Result::INVALID_ARGUMENT, "nrdcuSetProfiling: null context");
    ctx->profiling = enabled != 0;
    return 0;
}

// Synchronises the device, folds all pending per-dispatch timings into the per-pass totals and returns the number of passes seen
NRDCU_API uint32_t nrdcuResolveProfile(nrdcuContext* ctx) {
    if (!ctx) return 0;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (auto& t : ctx->pending) {
        float ms = 0.0f;
        if (cudaEventElapsedTime(&ms, t.start, t.stop) == cudaSuccess) {
            nrdcuContext::ProfileEntry* e = nullptr;
            for (auto& pe : ctx->profile)
                if (pe.name == t.name || !strcmp(pe.name, t.name)) e = &pe;
            if (!e) {
                ctx->profile.push_back({t.name});
                e = &ctx->profile.back();
            }
            e->totalMs += ms;
            e->count++;
        }
        ctx->freeEvents.push_back(t.start);
        ctx->freeEvents.push_back(t.stop);
    }
    ctx->pending.clear();
    return (uint32_t)ctx->profile.size();
}

NRDCU_API uint32_t nrdcuGetProfileEntry(nrdcuContext* ctx, uint32_t index, const char** name, double* totalMs, uint64_t* count) {
    if (!ctx || index >= ctx->profile.size()) return fail(Result::INVALID_ARGUMENT, "nrdcuGetProfileEntry: index out of range");
    if (name) *name = ctx->profile[index].name;
    if (totalMs) *totalMs = ctx->profile[index].totalMs;
    if (count) *count = ctx->profile[index].count;
    return 0;
}

NRDCU_API void nrdcuResetProfile(nrdcuContext* ctx) {
    if (ctx) ctx->profile.clear();
}

NRDCU_API uint32_t nrdcuSetHostResource(nrdcuContext* ctx, uint32_t resourceType, void* hostData, uint32_t width, uint32_t height, uint32_t pitchBytes, uint32_t format,
                                        int direction) {
    if (!ctx || resourceType >= (uint32_t)ResourceType::TRANSIENT_POOL) return fail(Result::INVALID_ARGUMENT, "nrdcuSetHostResource: bad argument");
    uint32_t bpp = bytesPerTexel(format);
    if (!bpp || (hostData && pitchBytes < width * bpp)) return fail(Result::INVALID_ARGUMENT, "nrdcuSetHostResource: bad format/pitch");
    if (!hostData) {
        // declaration for the host-frame path: the resource lives in nrdcuHostFrame blocks ( nrdcuHostFrameCreate lays them out )
        if (ctx->frameIn[0] || ctx->frameOut) return fail(Result::INVALID_ARGUMENT, "nrdcuSetHostResource: declare every host resource before the first nrdcuHostFrameCreate");
        nrdcuContext::FrameLayout& L = ctx->frameLayout[direction ? 1 : 0];
        for (const auto& sl : L.slots)
            if (sl.resourceType == resourceType) return fail(Result::INVALID_ARGUMENT, "nrdcuSetHostResource: resource %u declared twice", resourceType);
        nrdcuContext::FrameSlot sl;
        sl.resourceType = resourceType;
        sl.tex = {nullptr, width, height, (width * bpp + 255u) & ~255u, format};
        sl.offset = L.bytes;
        L.bytes += (size_t)sl.tex.pitchBytes * height;
        L.slots.push_back(sl);
        return 0;
    }
    cudaSetDevice(ctx->device);
    HostBinding& hb = ctx->hostBindings[resourceType];
    if (!hb.used || hb.device.width != width || hb.device.height != height || hb.device.format != format) {
        if (!allocTexture(ctx, format, width, height, hb.device, false)) return (uint32_t)Result::FAILURE;
    }
    hb.host = hostData;
    hb.hostPitch = pitchBytes;
    hb.direction = direction;
    hb.used = true;
    ctx->user[resourceType] = hb.device;
    return 0;
}

// Pipelined variant of nrdcuDenoiseHost: the H2D copy of frame i + 1 (own stream, second input buffer) and the D2H copy of frame
// i - 1 (own stream, from a staging copy of the outputs) overlap the kernels of frame i on `stream`. The call returns as soon as
// everything is enqueued; host outputs of this call are complete after nrdcuHostFlush( ctx, stream ) + a wait on `stream`.
// The host input buffers must stay untouched until then as well.
NRDCU_API uint32_t nrdcuDenoiseHostPipelined(nrdcuContext* ctx, const uint32_t* identifiers, uint32_t identifiersNum, void* stream) {
    if (!ctx) return fail(Result::INVALID_ARGUMENT, "nrdcuDenoiseHostPipelined: null context");
    cudaStream_t s = (cudaStream_t)stream;
    cudaSetDevice(ctx->device);
    ensurePipe(ctx);
    nrdcuContext::HostPipe& pp = ctx->pipe;
    const int b = (int)(pp.frame & 1);
    // second buffers on first use
    for (HostBinding& hb : ctx->hostBindings) {
        if (!hb.used || hb.deviceAlt.data) continue;
        if (!allocTexture(ctx, hb.device.format, hb.device.width, hb.device.height, hb.deviceAlt, false)) return (uint32_t)Result::FAILURE;
    }
    // inputs of this frame -> input buffer b, once the frame that last read buffer b (two calls ago) has consumed it
    if (pp.consumedValid[b]) cudaStreamWaitEvent(pp.h2d, pp.inConsumed[b], 0);
    for (size_t slot = 0; slot < (size_t)ResourceType::MAX_NUM; slot++) {
        HostBinding& hb = ctx->hostBindings[slot];
        if (!hb.used || hb.direction != 0) continue;
        const nrdcuTexture& dst = b ? hb.deviceAlt : hb.device;
        const uint32_t rowBytes = dst.width * bytesPerTexel(dst.format);
        cudaError_t e = cudaMemcpy2DAsync(dst.data, dst.pitchBytes, hb.host, hb.hostPitch, rowBytes, dst.height, cudaMemcpyHostToDevice, pp.h2d);
        if (e != cudaSuccess) return fail(Result::FAILURE, "H2D copy: %s", cudaGetErrorString(e));
        ctx->user[slot] = dst;
    }
    cudaEventRecord(pp.inReady[b], pp.h2d);
    cudaStreamWaitEvent(s, pp.inReady[b], 0);
    uint32_t rc = nrdcuDenoise(ctx, identifiers, identifiersNum, stream);
    if (rc != 0) return rc;
    cudaEventRecord(pp.inConsumed[b], s);
    pp.consumedValid[b] = true;
    // outputs: device-to-device into the staging copy (after the previous D2H has finished reading it), then D2H on its own stream
    if (pp.d2hValid) cudaStreamWaitEvent(s, pp.d2hDone, 0);
    for (HostBinding& hb : ctx->hostBindings) {
        if (!hb.used || hb.direction != 1) continue;
        const uint32_t rowBytes = hb.device.width * bytesPerTexel(hb.device.format);
        cudaError_t e = cudaMemcpy2DAsync(hb.deviceAlt.data, hb.deviceAlt.pitchBytes, hb.device.data, hb.device.pitchBytes, rowBytes, hb.device.height, cudaMemcpyDeviceToDevice, s);
        if (e != cudaSuccess) return fail(Result::FAILURE, "D2D copy: %s", cudaGetErrorString(e));
    }
    cudaEventRecord(pp.outCopied, s);
    cudaStreamWaitEvent(pp.d2h, pp.outCopied, 0);
    for (HostBinding& hb : ctx->hostBindings) {
        if (!hb.used || hb.direction != 1) continue;
        const uint32_t rowBytes = hb.device.width * bytesPerTexel(hb.device.format);
        cudaError_t e = cudaMemcpy2DAsync(hb.host, hb.hostPitch, hb.deviceAlt.data, hb.deviceAlt.pitchBytes, rowBytes, hb.device.height, cudaMemcpyDeviceToHost, pp.d2h);
        if (e != cudaSuccess) return fail(Result::FAILURE, "D2H copy: %s", cudaGetErrorString(e));
    }
    cudaEventRecord(pp.d2hDone, pp.d2h);
    pp.d2hValid = true;
    pp.frame++;
    return 0;
}

}  // extern "C"

// ---- host frames ----------------------------------------------------------------------------------------------------
struct nrdcuHostFrame {
    nrdcuContext* ctx = nullptr;
    int direction = 0;
    uint8_t* base = nullptr;
    size_t bytes = 0, mapped = 0;   // mapped != 0: mmap + cudaHostRegister ( NUMA-bound ); 0: cudaHostAlloc
    int numaNode = -1;
};

namespace {
// NUMA node of the GPU's PCIe root ( /sys/bus/pci/devices/<bus id>/numa_node ), -1 when the platform does not say
int numaNodeOfDevice(int device) {
    char busId[32] = {};
    if (cudaDeviceGetPCIBusId(busId, sizeof(busId), device) != cudaSuccess) return -1;
    for (char* c = busId; *c; c++) *c = (char)tolower(*c);
    char path[128];
    snprintf(path, sizeof(path), "/sys/bus/pci/devices/%s/numa_node", busId);
    FILE* f = fopen(path, "r");
    if (!f) return -1;
    int node = -1;
    if (fscanf(f, "%d", &node) != 1) node = -1;
    fclose(f);
    return node;
}

// Pinned host memory on the GPU's NUMA node: pages are bound with mbind( MPOL_PREFERRED ) before they are touched, then registered with
// CUDA. With all ranks of a node allocating from whatever node their thread happens to run on, the copies of the GPUs behind the other
// socket cross the inter-socket link and share one socket's DRAM ( VERDICT r1: e2e 2.75 -> 9.2 ms per frame from 1 to 8 GPUs ).
bool allocPinned(nrdcuHostFrame* f, int device) {
    const int node = numaNodeOfDevice(device);
    const size_t len = (f->bytes + (2u << 20) - 1) & ~(size_t)((2u << 20) - 1);
    if (node >= 0 && node < 1024) {
        void* p = mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        if (p != MAP_FAILED) {
            unsigned long mask[16] = {};
            mask[node / (8 * sizeof(unsigned long))] |= 1ul << (node % (8 * sizeof(unsigned long)));
            const long rc = syscall(SYS_mbind, p, len, 1 /* MPOL_PREFERRED */, mask, (unsigned long)(sizeof(mask) * 8), 0u);
            memset(p, 0, len);   // first touch under the policy
            if (cudaHostRegister(p, len, cudaHostRegisterPortable) == cudaSuccess) {
                f->base = (uint8_t*)p;
                f->mapped = len;
                f->numaNode = rc == 0 ? node : -1;
                return true;
            }
            cudaGetLastError();
            munmap(p, len);
        }
    }
    void* p = nullptr;
    if (cudaHostAlloc(&p, len, cudaHostAllocPortable) != cudaSuccess) return false;
    memset(p, 0, len);
    f->base = (uint8_t*)p;
    f->mapped = 0;
    f->numaNode = -1;
    return true;
}

void ensurePipe(nrdcuContext* ctx) {
    nrdcuContext::HostPipe& pp = ctx->pipe;
    if (pp.h2d) return;
    cudaStreamCreateWithFlags(&pp.h2d, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&pp.d2h, cudaStreamNonBlocking);
    for (int i = 0; i < 2; i++) {
        cudaEventCreateWithFlags(&pp.inReady[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&pp.inConsumed[i], cudaEventDisableTiming);
    }
    cudaEventCreateWithFlags(&pp.outCopied, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&pp.d2hDone, cudaEventDisableTiming);
}
}  // namespace

extern "C" {

NRDCU_API uint32_t nrdcuHostFrameCreate(nrdcuContext* ctx, int direction, nrdcuHostFrame** out) {
    if (!ctx || !out) return fail(Result::INVALID_ARGUMENT, "nrdcuHostFrameCreate: null argument");
    const int d = direction ? 1 : 0;
    const nrdcuContext::FrameLayout& L = ctx->frameLayout[d];
    if (L.slots.empty()) return fail(Result::INVALID_ARGUMENT, "nrdcuHostFrameCreate: no host resource declared for direction %d ( nrdcuSetHostResource with hostData = NULL )", d);
    cudaSetDevice(ctx->device);
    // device blocks on first use: two input buffers, or the block the kernels write + the staging copy the download reads
    uint8_t** blocks[2] = {d ? &ctx->frameOut : &ctx->frameIn[0], d ? &ctx->frameOutStage : &ctx->frameIn[1]};
    for (uint8_t** b : blocks) {
        if (*b) continue;
        void* p = nullptr;
        if (cudaMalloc(&p, L.bytes) != cudaSuccess) return fail(Result::FAILURE, "nrdcuHostFrameCreate: cudaMalloc of %zu bytes failed", L.bytes);
        cudaMemset(p, 0, L.bytes);
        ctx->allocations.push_back(p);
        *b = (uint8_t*)p;
    }
    nrdcuHostFrame* f = new nrdcuHostFrame();
    f->ctx = ctx;
    f->direction = d;
    f->bytes = L.bytes;
    if (!allocPinned(f, ctx->device)) {
        delete f;
        return fail(Result::FAILURE, "nrdcuHostFrameCreate: pinned host allocation of %zu bytes failed", L.bytes);
    }
    *out = f;
    return 0;
}

NRDCU_API uint32_t nrdcuHostFrameGetTexture(nrdcuHostFrame* frame, uint32_t resourceType, nrdcuTexture* outHostView) {
    if (!frame || !outHostView) return fail(Result::INVALID_ARGUMENT, "nrdcuHostFrameGetTexture: null argument");
    for (const auto& sl : frame->ctx->frameLayout[frame->direction].slots)
        if (sl.resourceType == resourceType) {
            *outHostView = sl.tex;
            outHostView->data = frame->base + sl.offset;
            return 0;
        }
    return fail(Result::INVALID_ARGUMENT, "nrdcuHostFrameGetTexture: resource %u is not part of this frame", resourceType);
}

NRDCU_API uint32_t nrdcuHostFrameGetInfo(nrdcuHostFrame* frame, uint64_t* bytes, int* numaNode) {
    if (!frame) return fail(Result::INVALID_ARGUMENT, "nrdcuHostFrameGetInfo: null frame");
    if (bytes) *bytes = frame->bytes;
    if (numaNode) *numaNode = frame->numaNode;
    return 0;
}

NRDCU_API void nrdcuHostFrameDestroy(nrdcuHostFrame* frame) {
    if (!frame) return;
    cudaSetDevice(frame->ctx->device);
    cudaDeviceSynchronize();   // a copy may still be reading / writing the block
    if (frame->mapped) {
        cudaHostUnregister(frame->base);
        munmap(frame->base, frame->mapped);
    } else if (frame->base)
        cudaFreeHost(frame->base);
    delete frame;
}

NRDCU_API uint32_t nrdcuDenoiseHostFrames(nrdcuContext* ctx, const uint32_t* identifiers, uint32_t identifiersNum, nrdcuHostFrame* inputs, nrdcuHostFrame* outputs, void* stream) {
    if (!ctx || !inputs || !outputs || inputs->ctx != ctx || outputs->ctx != ctx || inputs->direction != 0 || outputs->direction != 1)
        return fail(Result::INVALID_ARGUMENT, "nrdcuDenoiseHostFrames: needs an input and an output frame of this context");
    cudaStream_t s = (cudaStream_t)stream;
    cudaSetDevice(ctx->device);
    ensurePipe(ctx);
    nrdcuContext::HostPipe& pp = ctx->pipe;
    const int b = (int)(pp.frame & 1);
    // ONE upload: the whole input block, once the frame that last read device buffer b ( two calls ago ) has consumed it
    if (pp.consumedValid[b]) cudaStreamWaitEvent(pp.h2d, pp.inConsumed[b], 0);
    cudaError_t e = cudaMemcpyAsync(ctx->frameIn[b], inputs->base, inputs->bytes, cudaMemcpyHostToDevice, pp.h2d);
    if (e != cudaSuccess) return fail(Result::FAILURE, "H2D copy: %s", cudaGetErrorString(e));
    cudaEventRecord(pp.inReady[b], pp.h2d);
    cudaStreamWaitEvent(s, pp.inReady[b], 0);
    for (const auto& sl : ctx->frameLayout[0].slots) {
        ctx->user[sl.resourceType] = sl.tex;
        ctx->user[sl.resourceType].data = ctx->frameIn[b] + sl.offset;
    }
    for (const auto& sl : ctx->frameLayout[1].slots) {
        ctx->user[sl.resourceType] = sl.tex;
        ctx->user[sl.resourceType].data = ctx->frameOut + sl.offset;
    }
    uint32_t rc = nrdcuDenoise(ctx, identifiers, identifiersNum, stream);
    if (rc != 0) return rc;
    cudaEventRecord(pp.inConsumed[b], s);
    pp.consumedValid[b] = true;
    // outputs: one device copy into the staging block ( after the previous download has finished reading it ), then ONE download
    if (pp.d2hValid) cudaStreamWaitEvent(s, pp.d2hDone, 0);
    e = cudaMemcpyAsync(ctx->frameOutStage, ctx->frameOut, outputs->bytes, cudaMemcpyDeviceToDevice, s);
    if (e != cudaSuccess) return fail(Result::FAILURE, "D2D copy: %s", cudaGetErrorString(e));
    cudaEventRecord(pp.outCopied, s);
    cudaStreamWaitEvent(pp.d2h, pp.outCopied, 0);
    e = cudaMemcpyAsync(outputs->base, ctx->frameOutStage, outputs->bytes, cudaMemcpyDeviceToHost, pp.d2h);
    if (e != cudaSuccess) return fail(Result::FAILURE, "D2H copy: %s", cudaGetErrorString(e));
    cudaEventRecord(pp.d2hDone, pp.d2h);
    pp.d2hValid = true;
    pp.frame++;
    return 0;
}

// Makes `stream` wait for every copy the pipelined path has in flight (call before reading host outputs / stopping a timer)
NRDCU_API uint32_t nrdcuHostFlush(nrdcuContext* ctx, void* stream) {
    if (!ctx) return fail(Result::INVALID_ARGUMENT, "nrdcuHostFlush: null context");
    if (ctx->pipe.d2hValid) cudaStreamWaitEvent((cudaStream_t)stream, ctx->pipe.d2hDone, 0);
    return 0;
}

NRDCU_API uint32_t nrdcuDenoiseHost(nrdcuContext* ctx, const uint32_t* identifiers, uint32_t identifiersNum, void* stream) {
    if (!ctx) return fail(Result::INVALID_ARGUMENT, "nrdcuDenoiseHost: null context");
    cudaStream_t s = (cudaStream_t)stream;
    cudaSetDevice(ctx->device);
    if (ctx->pipe.d2hValid) cudaStreamWaitEvent(s, ctx->pipe.d2hDone, 0);  // a pipelined call may still be downloading
    for (size_t slot = 0; slot < (size_t)ResourceType::MAX_NUM; slot++) {
        HostBinding& hb = ctx->hostBindings[slot];
        if (!hb.used || hb.direction != 0) continue;
        uint32_t rowBytes = hb.device.width * bytesPerTexel(hb.device.format);
        cudaError_t e = cudaMemcpy2DAsync(hb.device.data, hb.device.pitchBytes, hb.host, hb.hostPitch, rowBytes, hb.device.height, cudaMemcpyHostToDevice, s);
        if (e != cudaSuccess) return fail(Result::FAILURE, "H2D copy: %s", cudaGetErrorString(e));
        ctx->user[slot] = hb.device;  // the pipelined path may have left the slot on the second input buffer
    }
    uint32_t rc = nrdcuDenoise(ctx, identifiers, identifiersNum, stream);
    if (rc != 0) return rc;
    for (HostBinding& hb : ctx->hostBindings) {
        if (!hb.used || hb.direction != 1) continue;
        uint32_t rowBytes = hb.device.width * bytesPerTexel(hb.device.format);
        cudaError_t e = cudaMemcpy2DAsync(hb.host, hb.hostPitch, hb.device.data, hb.device.pitchBytes, rowBytes, hb.device.height, cudaMemcpyDeviceToHost, s);
        if (e != cudaSuccess) return fail(Result::FAILURE, "D2H copy: %s", cudaGetErrorString(e));
    }
    return 0;
}

}  // extern "C"
