/* nrdcu.h — C ABI of the CUDA (sm_100a) executor for NRD dispatch streams.
 *
 * This is the piece that replaces the reference's NRI-based executor, nrd::Integration
 * (External/NRD/Integration/NRDIntegration.h:211-264, NRDIntegration.hpp:98-146 Recreate, :241-455 pool creation,
 * :522-620 Denoise, :723-890 per-dispatch binding + CmdDispatch): it owns the permanent/transient texture pools,
 * resolves every nrd::DispatchDesc binding to a pool or user texture and launches one hand-written CUDA kernel per
 * dispatch, keyed by PipelineDesc::shaderIdentifier (NRDDescs.h:452-453). Plain pointers and sizes only — no torch,
 * no C++ types — so the reference-side binding is a dozen lines.
 *
 * Texture contract ("bit-identical layout"): pitch-linear 2D arrays in device memory, texel encodings exactly the
 * nrd::Format values (NRDDescs.h:264-322): IN_VIEWZ R32_SFLOAT, IN_MV RGBA16_SFLOAT, IN_NORMAL_ROUGHNESS
 * R10_G10_B10_A2_UNORM, IN/OUT_*_RADIANCE_HITDIST RGBA16_SFLOAT (YCoCg + normalised hit distance), IN_PENUMBRA
 * R16_SFLOAT, OUT_SHADOW_TRANSLUCENCY R8_UNORM. `data` and `pitchBytes` must be multiples of the texel size (pools owned by the executor use a 256-byte row pitch).
 *
 * Every function returns an nrd::Result value as uint32_t (0 = SUCCESS, 1 = FAILURE, 2 = INVALID_ARGUMENT,
 * 3 = UNSUPPORTED); nrdcuGetLastError() describes the last non-success on the calling thread.
 * There is no CPU fallback: without a CUDA device every launching call returns FAILURE.
 */
#ifndef NRDCU_H
#define NRDCU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef NRDCU_API
#    define NRDCU_API __attribute__((visibility("default")))
#endif

typedef struct nrdcuContext nrdcuContext; /* opaque: nrd::Instance + pools + stream state */

/* One 2D texture in device memory. `format` is an nrd::Format value. Same layout as the oracle's view. */
typedef struct nrdcuTexture {
    void* data;
    uint32_t width, height;
    uint32_t pitchBytes;
    uint32_t format;
} nrdcuTexture;

/* Flags of nrdcuCreate / nrdcuDispatch */
enum {
    NRDCU_FLAG_QUAD_INTRINSICS = 1u << 0, /* replay SM6.0 quad smoothing (NRD_SUPPORTS_QUAD_INTRINSICS=1, the reference default) */
    NRDCU_FLAG_CUDA_GRAPH = 1u << 1,      /* nrdcuDenoise keeps each distinct frame ( the chain of kernels and the memory it binds: odd / even ping-pong, the
                                             first frame with its clears ) as an instantiated CUDA graph; later frames of the same shape only patch the kernel
                                             nodes' parameters ( constants ) and launch the graph once. Pays where the frame is launch-bound ( small frames );
                                             ignored with strips over peer memory, profiling, row ranges and per-dispatch callbacks */
    NRDCU_FLAG_ROBUST_MIRROR_TEST = 1u << 2, /* DEBUG: spatial taps use "left the screen" instead of the reference's bit-fragile any(uv != MirrorUv(uv)) */
    NRDCU_FLAG_NO_SEAM_OVERLAP = 1u << 4,    /* strips over peer memory: do NOT split each pass into "seam rows first, interior second" ( the push of the seam rows then
                                                follows the whole pass instead of overlapping its interior ); for A/B timing */
    NRDCU_FLAG_PROBE_MIRROR = 1u << 3,       /* DEBUG: the REBLUR_DIFFUSE_SPECULAR spatial passes count their taps and how many took the "mirrored" branch
                                                of REBLUR_Common_SpatialFilter.hlsli:198 ( read with nrdcuGetMirrorProbe ) */
};
#define NRDCU_DEFAULT_FLAGS (NRDCU_FLAG_QUAD_INTRINSICS)

/* ---- low level: one pass ------------------------------------------------------------------------------------
 * Replaces one NRDIntegration::_Dispatch (NRDIntegration.hpp:723-890). `textures` follow DispatchDesc::resources
 * order (inputs then outputs); `constants` are DispatchDesc::constantBufferData (host memory, copied into the launch).
 * `stream` is a cudaStream_t. Asynchronous. */
NRDCU_API uint32_t nrdcuDispatch(const char* shaderIdentifier, const void* constants, uint32_t constantsSize, const nrdcuTexture* textures,
                                 uint32_t texturesNum, uint32_t flags, void* stream);
/* Same, restricted to the rect rows [rowBegin, rowEnd) (rowBegin a multiple of 16; rowEnd is clamped to the rect). Every texture is
 * still the whole frame: a strip reads rows outside its range (the halo) and writes only rows inside it. Every REBLUR, RELAX and SIGMA pass takes a
 * range ( SIGMA's two 1/16-resolution tile passes ignore it and cover the frame ); REFERENCE and the validation overlays are whole-frame only. */
NRDCU_API uint32_t nrdcuDispatchRows(const char* shaderIdentifier, const void* constants, uint32_t constantsSize, const nrdcuTexture* textures,
                                     uint32_t texturesNum, uint32_t flags, void* stream, uint32_t rowBegin, uint32_t rowEnd);

/* ---- instance level ------------------------------------------------------------------------------------------
 * nrdcuCreate           == Integration::Recreate: nrd::CreateInstance + pool allocation at `resourceWidth x Height`
 *                          (instanceCreationDesc is a `const nrd::InstanceCreationDesc*`)
 * nrdcuSetCommonSettings / nrdcuSetDenoiserSettings == the nrd:: calls of the same name (settings structs by pointer)
 * nrdcuSetResource      == one slot of ResourceSnapshot (NRDIntegration.h:121-167): user texture for a ResourceType
 * nrdcuDenoise          == Integration::Denoise: GetComputeDispatches + one kernel launch per dispatch on `stream`
 * nrdcuGetPoolTexture   -- debug / parity tap on a pool texture (isPermanent != 0 -> permanent pool)
 * nrdcuGetInstance      -- the underlying nrd::Instance*, for callers that want the descriptor API as well */
NRDCU_API uint32_t nrdcuCreate(const void* instanceCreationDesc, uint16_t resourceWidth, uint16_t resourceHeight, int device, uint32_t flags, nrdcuContext** out);
NRDCU_API void nrdcuDestroy(nrdcuContext* ctx);
NRDCU_API uint32_t nrdcuSetCommonSettings(nrdcuContext* ctx, const void* commonSettings);
NRDCU_API uint32_t nrdcuSetDenoiserSettings(nrdcuContext* ctx, uint32_t identifier, const void* denoiserSettings);
NRDCU_API uint32_t nrdcuSetResource(nrdcuContext* ctx, uint32_t resourceType, const nrdcuTexture* texture);
NRDCU_API uint32_t nrdcuDenoise(nrdcuContext* ctx, const uint32_t* identifiers, uint32_t identifiersNum, void* stream);
/* One frame tiled over several GPUs as horizontal strips (there is no counterpart in NRDIntegration: the reference runs one queue on
 * one device). Like nrdcuDenoise, but every pass computes only rows [rowBegin, rowEnd); after each dispatch has been enqueued
 * `afterDispatch` (may be NULL) is called with the dispatch's resolved textures and a flag per texture telling whether the pass wrote
 * it, so the caller can trade the seam rows of those textures with the neighbouring strips (nrd_sample_b200/tiling.py: NCCL
 * send/recv on the same stream) before the next pass reads them. */
typedef void (*nrdcuDispatchCallback)(void* userArg, uint32_t dispatchIndex, const char* passName, const nrdcuTexture* textures, const uint8_t* isStorage,
                                      uint32_t texturesNum);
NRDCU_API uint32_t nrdcuDenoiseRows(nrdcuContext* ctx, const uint32_t* identifiers, uint32_t identifiersNum, void* stream, uint32_t rowBegin, uint32_t rowEnd,
                                    nrdcuDispatchCallback afterDispatch, void* userArg);
NRDCU_API uint32_t nrdcuGetPoolTexture(nrdcuContext* ctx, int isPermanent, uint32_t index, nrdcuTexture* out);
NRDCU_API void* nrdcuGetInstance(nrdcuContext* ctx);

/* ---- strips over peer memory -----------------------------------------------------------------------------------
 * The B200-native seam exchange for nrdcuDenoiseRows: one process per GPU, neighbouring strips map each other's textures
 * through CUDA IPC and every pass pushes its seam rows straight into the neighbours' copies over NVLink, then raises a flag
 * the neighbour spins on before its next pass (csrc/kernels/peer_halo.cu) — no host round trip, no collective library.
 *   1. every rank: nrdcuCreate; nrdcuAllocSharedTexture for each user texture the denoiser WRITES (OUT_*), bound with
 *      nrdcuSetResource like any other texture; optional nrdcuTileSetHalo rules (apron rows per pass / binding, default 64)
 *   2. every rank: nrdcuTileExport -> a blob of IPC handles; ranks trade blobs (torch.distributed / MPI / a pipe: plumbing)
 *   3. every rank: nrdcuTileAttach(blob of the rank above, blob of the rank below)  (NULL at the top / bottom of the frame)
 *   4. per frame, in lockstep: nrdcuDenoiseRows(ctx, ..., rowBegin, rowEnd, NULL, NULL)
 * nrdcuTileGetStatus: bytes pushed so far and a non-zero error if a wait for a neighbour timed out (~4 s).
 * Row ranges work for every denoiser ( REBLUR, RELAX, SIGMA; REFERENCE has none ): SIGMA's two tile passes ignore the range and cover the whole frame. */
NRDCU_API uint32_t nrdcuAllocSharedTexture(nrdcuContext* ctx, uint32_t format, uint32_t width, uint32_t height, nrdcuTexture* out);
NRDCU_API uint32_t nrdcuTileExportSize(nrdcuContext* ctx);
NRDCU_API uint32_t nrdcuTileExport(nrdcuContext* ctx, void* blob, uint32_t blobSize);
NRDCU_API uint32_t nrdcuTileAttach(nrdcuContext* ctx, const void* blobAbove, const void* blobBelow, uint32_t blobSize);
NRDCU_API uint32_t nrdcuTileSetHalo(nrdcuContext* ctx, const char* passName /* NULL: set the default */, uint32_t binding, uint32_t rows);
NRDCU_API uint32_t nrdcuTileGetStatus(nrdcuContext* ctx, uint64_t* bytesPushed, uint32_t* error);
/* The temporal passes fetch history at pixel + motion, and a strip only holds `halo` rows of its neighbours: vertical motion of more than
 * *boundRows ( = halo - 2 ) rows per frame reads rows that were never delivered. nrdcuDenoiseRows checks every strip frame on the device ( 2D / 2.5D motion
 * vectors; one pass over the strip's IN_MV ): *worstExcessRows is the largest overshoot seen so far, 0 = every fetch stayed inside the apron. A renderer
 * that sees it non-zero needs a taller halo ( nrdcuTileSetHalo( ctx, NULL, 0, rows ) before nrdcuTileExport ) for that camera speed. */
NRDCU_API uint32_t nrdcuTileGetMotionBound(nrdcuContext* ctx, uint32_t* boundRows, uint32_t* worstExcessRows);

/* Host-buffer convenience used by plugin-style callers (the `e2e` path of bench.py): uploads the user inputs that
 * were registered with nrdcuSetHostResource from pinned/pageable host memory, runs nrdcuDenoise, downloads the outputs.
 * direction: 0 = input (H2D before the frame), 1 = output (D2H after the frame). Device staging is owned by ctx. */
NRDCU_API uint32_t nrdcuSetHostResource(nrdcuContext* ctx, uint32_t resourceType, void* hostData, uint32_t width, uint32_t height, uint32_t pitchBytes,
                                        uint32_t format, int direction);
NRDCU_API uint32_t nrdcuDenoiseHost(nrdcuContext* ctx, const uint32_t* identifiers, uint32_t identifiersNum, void* stream);
/* Same work, pipelined over PCIe: the upload of the NEXT call's inputs (second device buffer, own copy stream) and the download of the
 * PREVIOUS call's outputs (from a staging copy, own copy stream) overlap this call's kernels on `stream`. Returns once everything is
 * enqueued. Host outputs are complete, and host inputs may be reused, after nrdcuHostFlush( ctx, stream ) followed by a wait on `stream`. */
NRDCU_API uint32_t nrdcuDenoiseHostPipelined(nrdcuContext* ctx, const uint32_t* identifiers, uint32_t identifiersNum, void* stream);
NRDCU_API uint32_t nrdcuHostFlush(nrdcuContext* ctx, void* stream);

/* Host frames: the plugin path without per-texture copies. A host frame is ONE pinned host block holding every host resource of one
 * direction ( declared with nrdcuSetHostResource( ..., hostData = NULL, ... ) ), laid out exactly like the executor's device block
 * ( 256-byte row pitch ), allocated on the NUMA node the GPU hangs off ( mbind + cudaHostRegister; falls back to cudaHostAlloc ). The
 * renderer writes its G-buffer / radiance straight into the views nrdcuHostFrameGetTexture returns ( `data` is a HOST pointer there ) and
 * reads the denoised outputs from an output frame, so a frame costs ONE cudaMemcpyAsync up and ONE down, pipelined like
 * nrdcuDenoiseHostPipelined ( upload of call i + 1 and download of call i - 1 overlap the kernels of call i; nrdcuHostFlush as there ).
 * Any number of frames may exist per direction ( e.g. a ring the renderer fills ahead ). */
typedef struct nrdcuHostFrame nrdcuHostFrame;
NRDCU_API uint32_t nrdcuHostFrameCreate(nrdcuContext* ctx, int direction, nrdcuHostFrame** out);
NRDCU_API uint32_t nrdcuHostFrameGetTexture(nrdcuHostFrame* frame, uint32_t resourceType, nrdcuTexture* outHostView);
NRDCU_API uint32_t nrdcuHostFrameGetInfo(nrdcuHostFrame* frame, uint64_t* bytes, int* numaNode /* -1: unknown / not bound */);
NRDCU_API void nrdcuHostFrameDestroy(nrdcuHostFrame* frame);
NRDCU_API uint32_t nrdcuDenoiseHostFrames(nrdcuContext* ctx, const uint32_t* identifiers, uint32_t identifiersNum, nrdcuHostFrame* inputs, nrdcuHostFrame* outputs,
                                          void* stream);

/* ---- per-pass timing -----------------------------------------------------------------------------------------
 * With profiling on, nrdcuDenoise brackets every dispatch with CUDA events on `stream`; nrdcuResolveProfile
 * synchronises and accumulates them per pass name (DispatchDesc::name). Used by bench.py for the live roofline. */
NRDCU_API uint32_t nrdcuSetProfiling(nrdcuContext* ctx, int enabled);
NRDCU_API uint32_t nrdcuResolveProfile(nrdcuContext* ctx);
NRDCU_API uint32_t nrdcuGetProfileEntry(nrdcuContext* ctx, uint32_t index, const char** name, double* totalMs, uint64_t* count);
NRDCU_API void nrdcuResetProfile(nrdcuContext* ctx);

/* ---- front end / back end on the device ( SURVEY.md 8(f).2 ) ---------------------------------------------------------
 * The application-side helpers of NRD.hlsli as CUDA: include/nrd_frontend.cuh holds the functions ( host + device ), these entry points run
 * them over whole frames for renderers that keep fp32 results in linear device buffers ( tightly packed, width * height texels ).
 * Replaces the packing at NRDSample Shaders/TraceOpaque.cs.hlsl:738-757 and the unpacking at Shaders/Composition.cs.hlsl:85-118.
 *   mode 0 = REBLUR ( YCoCg radiance, hit distance normalised with ReblurSettings::hitDistanceParameters { A, B, C } ), 1 = RELAX.
 * Return values are nrd::Result; nrdcuFrontEndGetLastError describes the last failure of these calls. All calls are asynchronous on `stream`
 * except nrdcuFrontEndProbe. */
NRDCU_API uint32_t nrdcuFrontEndPackNormalRoughness(const float* normalRoughness /* float4: N.xyz, linear roughness */, const float* materialID /* 0..3, may be NULL */,
                                                    const nrdcuTexture* outNormalRoughness /* R10_G10_B10_A2_UNORM */, void* stream);
NRDCU_API uint32_t nrdcuFrontEndPackRadianceHitDist(uint32_t mode, const float* radianceHitDist /* float4: linear radiance, hit distance */, const nrdcuTexture* viewZ /* R32_SFLOAT, REBLUR */,
                                                    const nrdcuTexture* normalRoughness /* specular lobe of REBLUR */, uint32_t isSpecular, const float* hitDistParams3,
                                                    const nrdcuTexture* out /* RGBA16_SFLOAT */, void* stream);
NRDCU_API uint32_t nrdcuBackEndUnpackRadiance(uint32_t mode, const nrdcuTexture* in /* RGBA16_SFLOAT */, float* outRadiance /* float4: linear radiance, .w as stored */, void* stream);
/* Verification hook: nrd_sample_b200/csrc/frontend_probe.inl ( every function of nrd_frontend.cuh ) over n columns; in6 / out21 are HOST arrays of DEVICE pointers to n float4 each. */
NRDCU_API uint32_t nrdcuFrontEndProbe(const float* const* in6, float* const* out21, uint32_t n, void* stream);
NRDCU_API const char* nrdcuFrontEndGetLastError(void);

/* ---- introspection ------------------------------------------------------------------------------------------- */
NRDCU_API const char* nrdcuGetLastError(void);
NRDCU_API uint64_t nrdcuGetLaunchCount(void);           /* kernels launched by this library since load (all contexts) */
NRDCU_API uint64_t nrdcuGetPoolBytes(nrdcuContext* ctx); /* device bytes held by the pools (README memory table analogue) */
/* The split NRDIntegration.h:266-277 reports: permanent pool ( history, survives the frame ), transient pool ( aliasable between denoisers and with the
 * application's own per-frame memory ), plus what this executor keeps for itself ( REBLUR's geometry plane ). Any pointer may be NULL. */
NRDCU_API uint32_t nrdcuGetMemoryUsage(nrdcuContext* ctx, uint64_t* persistentBytes, uint64_t* aliasableBytes, uint64_t* privateBytes);
/* NRDCU_FLAG_CUDA_GRAPH bookkeeping: frames captured into a new graph, frames replayed from a cached one, graphs currently cached */
NRDCU_API uint32_t nrdcuGetGraphStats(nrdcuContext* ctx, uint64_t* captures, uint64_t* replays, uint32_t* cached);
/* NRDCU_FLAG_PROBE_MIRROR counters of the current device, 14 values: out[0] = taps, out[1] = taps whose weight took the "mirrored" branch, then the same pair
 * per ( pass, lobe ) at out[2 + 2 * slot], slot = pass * 2 + lobe ( pass 0 pre-pass / 1 blur / 2 post-blur; lobe 0 diffuse / 1 specular ); reset != 0 clears them */
NRDCU_API uint32_t nrdcuGetMirrorProbe(uint64_t* out, int reset);

#ifdef __cplusplus
}
#endif
#endif /* NRDCU_H */
