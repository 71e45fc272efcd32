// nrd_b200.h — descriptor-level C ABI of the B200-native NRD replacement.
//
// Layout- and symbol-compatible restatement of the reference's public interface
// (External/NRD/Include/NRD.h:60-79, NRDDescs.h:18-529, NRDSettings.h:18-488, v4.17): the nine
// `extern "C"` entry points below are exactly the ones NRDIntegration.hpp / NRDSample.cpp bind,
// every POD has the same field order, size and defaults, every enum the same numeric values.
// A caller compiled against the reference headers can link this library unchanged; the static
// asserts at the bottom pin the sizes measured from the reference build (oracle/_ref).
//
// What differs from the reference: PipelineDesc bytecode pointers are always null (there is no
// DXIL/SPIR-V here) and `shaderIdentifier` is the key the CUDA executor (include/nrdcu.h)
// dispatches on — the use NRDDescs.h:452-453 reserves for custom integrations.
#pragma once

#include <cstddef>
#include <cstdint>

#define NRD_VERSION_MAJOR 4
#define NRD_VERSION_MINOR 17
#define NRD_VERSION_BUILD 4
#define NRD_VERSION_DATE "2 May 2026"
#define NRD_DESCS_VERSION_MAJOR 4
#define NRD_DESCS_VERSION_MINOR 17
#define NRD_SETTINGS_VERSION_MAJOR 4
#define NRD_SETTINGS_VERSION_MINOR 17

#if defined(_WIN32)
#    define NRD_CALL __stdcall
#else
#    define NRD_CALL
#endif
#ifndef NRD_API
#    define NRD_API extern "C"
#endif

namespace nrd {

typedef uint32_t Identifier;
struct Instance;  // opaque

// ---------------------------------------------------------------------------------------------
// Enums (NRDDescs.h:24-322, NRDSettings.h:43-81) — numeric values are ABI
// ---------------------------------------------------------------------------------------------
enum class Result : uint32_t { SUCCESS, FAILURE, INVALID_ARGUMENT, UNSUPPORTED, NON_UNIQUE_IDENTIFIER, MAX_NUM };

enum class ResourceType : uint32_t {
    // guides
    IN_MV, IN_NORMAL_ROUGHNESS, IN_VIEWZ,
    // optional guides
    IN_DIFF_CONFIDENCE, IN_SPEC_CONFIDENCE, IN_DISOCCLUSION_THRESHOLD_MIX,
    // noisy signals
    IN_DIFF_RADIANCE_HITDIST, IN_SPEC_RADIANCE_HITDIST, IN_DIFF_HITDIST, IN_SPEC_HITDIST, IN_DIFF_DIRECTION_HITDIST,
    IN_DIFF_SH0, IN_DIFF_SH1, IN_SPEC_SH0, IN_SPEC_SH1, IN_PENUMBRA, IN_TRANSLUCENCY, IN_SIGNAL,
    // denoised signals
    OUT_DIFF_RADIANCE_HITDIST, OUT_SPEC_RADIANCE_HITDIST, OUT_DIFF_SH0, OUT_DIFF_SH1, OUT_SPEC_SH0, OUT_SPEC_SH1,
    OUT_DIFF_HITDIST, OUT_SPEC_HITDIST, OUT_DIFF_DIRECTION_HITDIST, OUT_SHADOW_TRANSLUCENCY, OUT_SIGNAL, OUT_VALIDATION,
    // pools owned by the executor
    TRANSIENT_POOL, PERMANENT_POOL,
    MAX_NUM,
};

enum class Denoiser : uint32_t {
    REBLUR_DIFFUSE, REBLUR_DIFFUSE_OCCLUSION, REBLUR_DIFFUSE_SH,
    REBLUR_SPECULAR, REBLUR_SPECULAR_OCCLUSION, REBLUR_SPECULAR_SH,
    REBLUR_DIFFUSE_SPECULAR, REBLUR_DIFFUSE_SPECULAR_OCCLUSION, REBLUR_DIFFUSE_SPECULAR_SH,
    REBLUR_DIFFUSE_DIRECTIONAL_OCCLUSION,
    RELAX_DIFFUSE, RELAX_DIFFUSE_SH, RELAX_SPECULAR, RELAX_SPECULAR_SH, RELAX_DIFFUSE_SPECULAR, RELAX_DIFFUSE_SPECULAR_SH,
    SIGMA_SHADOW, SIGMA_SHADOW_TRANSLUCENCY,
    REFERENCE,
    MAX_NUM
};

enum class Format : uint32_t {
    R8_UNORM, R8_SNORM, R8_UINT, R8_SINT,
    RG8_UNORM, RG8_SNORM, RG8_UINT, RG8_SINT,
    RGBA8_UNORM, RGBA8_SNORM, RGBA8_UINT, RGBA8_SINT, RGBA8_SRGB,
    R16_UNORM, R16_SNORM, R16_UINT, R16_SINT, R16_SFLOAT,
    RG16_UNORM, RG16_SNORM, RG16_UINT, RG16_SINT, RG16_SFLOAT,
    RGBA16_UNORM, RGBA16_SNORM, RGBA16_UINT, RGBA16_SINT, RGBA16_SFLOAT,
    R32_UINT, R32_SINT, R32_SFLOAT,
    RG32_UINT, RG32_SINT, RG32_SFLOAT,
    RGB32_UINT, RGB32_SINT, RGB32_SFLOAT,
    RGBA32_UINT, RGBA32_SINT, RGBA32_SFLOAT,
    R10_G10_B10_A2_UNORM, R10_G10_B10_A2_UINT, R11_G11_B10_UFLOAT, R9_G9_B9_E5_UFLOAT,
    MAX_NUM
};

enum class DescriptorType : uint32_t { TEXTURE, STORAGE_TEXTURE, MAX_NUM };
enum class Sampler : uint32_t { NEAREST_CLAMP, LINEAR_CLAMP, MAX_NUM };
enum class NormalEncoding : uint8_t { RGBA8_UNORM, RGBA8_SNORM, R10_G10_B10_A2_UNORM, RGBA16_UNORM, RGBA16_SNORM, MAX_NUM };
enum class RoughnessEncoding : uint8_t { SQ_LINEAR, LINEAR, SQRT_LINEAR, MAX_NUM };
enum class CheckerboardMode : uint8_t { OFF, BLACK, WHITE, MAX_NUM };
enum class AccumulationMode : uint8_t { CONTINUE, RESTART, CLEAR_AND_RESTART, MAX_NUM };
enum class HitDistanceReconstructionMode : uint8_t { OFF, AREA_3X3, AREA_5X5, MAX_NUM };

// ---------------------------------------------------------------------------------------------
// Descriptors (NRDDescs.h:378-529)
// ---------------------------------------------------------------------------------------------
struct AllocationCallbacks {
    void*(NRD_CALL* Allocate)(void* userArg, size_t size, size_t alignment);
    void*(NRD_CALL* Reallocate)(void* userArg, void* memory, size_t size, size_t alignment);
    void(NRD_CALL* Free)(void* userArg, void* memory);
    void* userArg;
};

struct SPIRVBindingOffsets {
    uint32_t samplerOffset, textureOffset, constantBufferOffset, storageTextureAndBufferOffset;
};

struct LibraryDesc {
    SPIRVBindingOffsets spirvBindingOffsets;
    const Denoiser* supportedDenoisers;
    uint32_t supportedDenoisersNum;
    uint8_t versionMajor, versionMinor, versionBuild;
    NormalEncoding normalEncoding;
    RoughnessEncoding roughnessEncoding;
};

struct DenoiserDesc {
    Identifier identifier;
    Denoiser denoiser;
};

struct InstanceCreationDesc {
    AllocationCallbacks allocationCallbacks;
    const DenoiserDesc* denoisers;
    uint32_t denoisersNum;
};

struct TextureDesc {
    Format format;
    uint16_t downsampleFactor;
};

struct ResourceDesc {
    DescriptorType descriptorType;
    ResourceType type;
    uint16_t indexInPool;
};

struct ResourceRangeDesc {
    DescriptorType descriptorType;
    uint32_t descriptorsNum;
};

struct ComputeShaderDesc {
    const void* bytecode;  // always null in this library
    uint64_t size;
};

struct PipelineDesc {
    ComputeShaderDesc computeShaderDXBC, computeShaderDXIL, computeShaderSPIRV;
    const ResourceRangeDesc* resourceRanges;
    uint32_t resourceRangesNum;
    bool hasConstantData;
    char shaderIdentifier[256];  // "File.cs.hlsl|MACRO=VAL|..." — the CUDA kernel key
};

struct DescriptorPoolDesc {
    uint32_t perSetTexturesMaxNum, perSetStorageTexturesMaxNum, totalTexturesNum, totalStorageTexturesNum, setsMaxNum;
};

struct InstanceDesc {
    uint32_t constantBufferAndSamplersSpaceIndex, resourcesSpaceIndex, constantBufferRegisterIndex;
    uint32_t samplersBaseRegisterIndex, resourcesBaseRegisterIndex;
    uint32_t constantBufferMaxDataSize;
    const Sampler* samplers;
    uint32_t samplersNum;
    const char* shaderEntryPoint;
    const PipelineDesc* pipelines;
    uint32_t pipelinesNum;
    const TextureDesc* permanentPool;
    uint32_t permanentPoolSize;
    const TextureDesc* transientPool;
    uint32_t transientPoolSize;
    DescriptorPoolDesc descriptorPoolDesc;
};

struct DispatchDesc {
    const char* name;
    Identifier identifier;
    const ResourceDesc* resources;  // inputs first, then outputs, in shader binding order
    uint32_t resourcesNum;
    const uint8_t* constantBufferData;
    uint32_t constantBufferDataSize;
    bool constantBufferDataMatchesPreviousDispatch;
    uint16_t pipelineIndex;
    uint16_t gridWidth, gridHeight;
};

// ---------------------------------------------------------------------------------------------
// Settings (NRDSettings.h:84-488) — defaults are part of the contract
// ---------------------------------------------------------------------------------------------
inline uint32_t GetMaxAccumulatedFrameNum(float accumulationTime, float fps) { return (uint32_t)(accumulationTime * fps + 0.5f); }

struct CommonSettings {
    float viewToClipMatrix[16] = {};
    float viewToClipMatrixPrev[16] = {};
    float worldToViewMatrix[16] = {};
    float worldToViewMatrixPrev[16] = {};
    float worldPrevToWorldMatrix[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    float motionVectorScale[3] = {1.0f, 1.0f, 0.0f};
    float cameraJitter[2] = {};
    float cameraJitterPrev[2] = {};
    uint16_t resourceSize[2] = {};
    uint16_t resourceSizePrev[2] = {};
    uint16_t rectSize[2] = {};
    uint16_t rectSizePrev[2] = {};
    float viewZScale = 1.0f;
    float timeDeltaBetweenFrames = 0.0f;
    float denoisingRange = 500000.0f;
    float disocclusionThreshold = 0.01f;
    float disocclusionThresholdAlternate = 0.05f;
    float cameraAttachedReflectionMaterialID = 999.0f;
    float strandMaterialID = 999.0f;
    float historyFixAlternatePixelStrideMaterialID = 999.0f;
    float strandThickness = 80e-6f;
    float splitScreen = 0.0f;
    uint16_t printfAt[2] = {9999, 9999};
    float debug = 0.0f;
    uint32_t rectOrigin[2] = {};
    uint32_t frameIndex = 0;
    AccumulationMode accumulationMode = AccumulationMode::CONTINUE;
    bool isMotionVectorInWorldSpace = false;
    bool isHistoryConfidenceAvailable = false;
    bool isDisocclusionThresholdMixAvailable = false;
    bool enableValidation = false;
};

const uint32_t REBLUR_MAX_HISTORY_FRAME_NUM = 63;
const float REBLUR_DEFAULT_ACCUMULATION_TIME = 0.5f;

struct ReblurHitDistanceParameters { float A = 3.0f, B = 0.1f, C = 20.0f; };
struct ReblurAntilagSettings { float luminanceSigmaScale = 2.0f, luminanceSensitivity = 3.0f; };
struct ReblurResponsiveAccumulationSettings { float roughnessThreshold = 0.0f; uint32_t minAccumulatedFrameNum = 3; };
struct ReblurConvergenceSettings { float s = 1.0f, b = 0.2f, p = 0.8f; };

struct ReblurSettings {
    ReblurHitDistanceParameters hitDistanceParameters = {};
    ReblurAntilagSettings antilagSettings = {};
    ReblurResponsiveAccumulationSettings responsiveAccumulationSettings = {};
    ReblurConvergenceSettings convergenceSettings = {};
    uint32_t maxAccumulatedFrameNum = 30;
    uint32_t maxFastAccumulatedFrameNum = 6;
    uint32_t maxStabilizedFrameNum = REBLUR_MAX_HISTORY_FRAME_NUM;
    uint32_t historyFixFrameNum = 3;
    uint32_t historyFixBasePixelStride = 14;
    uint32_t historyFixAlternatePixelStride = 14;
    float fastHistoryClampingSigmaScale = 2.0f;
    float diffusePrepassBlurRadius = 30.0f;
    float specularPrepassBlurRadius = 50.0f;
    float minHitDistanceWeight = 0.1f;
    float minBlurRadius = 1.0f;
    float maxBlurRadius = 30.0f;
    float lobeAngleFraction = 0.15f;
    float roughnessFraction = 0.15f;
    float planeDistanceSensitivity = 0.02f;
    float fireflySuppressorMinRelativeScale = 2.0f;
    float minMaterialForDiffuse = 4.0f;
    float minMaterialForSpecular = 4.0f;
    CheckerboardMode checkerboardMode = CheckerboardMode::OFF;
    HitDistanceReconstructionMode hitDistanceReconstructionMode = HitDistanceReconstructionMode::OFF;
    bool enableAntiFirefly = true;
    bool usePrepassOnlyForSpecularMotionEstimation = false;
    bool returnHistoryLengthInsteadOfOcclusion = false;
};

const uint32_t RELAX_MAX_HISTORY_FRAME_NUM = 255;
const float RELAX_DEFAULT_ACCUMULATION_TIME = 0.5f;

struct RelaxAntilagSettings { float accelerationAmount = 0.3f, spatialSigmaScale = 4.5f, temporalSigmaScale = 0.5f, resetAmount = 0.5f; };

struct RelaxSettings {
    RelaxAntilagSettings antilagSettings = {};
    uint32_t diffuseMaxAccumulatedFrameNum = 30;
    uint32_t specularMaxAccumulatedFrameNum = 30;
    uint32_t diffuseMaxFastAccumulatedFrameNum = 6;
    uint32_t specularMaxFastAccumulatedFrameNum = 6;
    uint32_t historyFixFrameNum = 3;
    uint32_t historyFixBasePixelStride = 14;
    uint32_t historyFixAlternatePixelStride = 14;
    float historyFixEdgeStoppingNormalPower = 8.0f;
    float fastHistoryClampingSigmaScale = 2.0f;
    float diffusePrepassBlurRadius = 30.0f;
    float specularPrepassBlurRadius = 50.0f;
    float minHitDistanceWeight = 0.1f;
    uint32_t spatialVarianceEstimationHistoryThreshold = 3;
    float diffusePhiLuminance = 2.0f;
    float specularPhiLuminance = 1.0f;
    float lobeAngleFraction = 0.5f;
    float roughnessFraction = 0.15f;
    float specularVarianceBoost = 0.0f;
    float specularLobeAngleSlack = 0.15f;
    uint32_t atrousIterationNum = 5;
    float diffuseMinLuminanceWeight = 0.0f;
    float specularMinLuminanceWeight = 0.0f;
    float depthThreshold = 0.003f;
    float confidenceDrivenRelaxationMultiplier = 0.0f;
    float confidenceDrivenLuminanceEdgeStoppingRelaxation = 0.0f;
    float confidenceDrivenNormalEdgeStoppingRelaxation = 0.0f;
    float luminanceEdgeStoppingRelaxation = 0.5f;
    float normalEdgeStoppingRelaxation = 0.3f;
    float roughnessEdgeStoppingRelaxation = 1.0f;
    CheckerboardMode checkerboardMode = CheckerboardMode::OFF;
    HitDistanceReconstructionMode hitDistanceReconstructionMode = HitDistanceReconstructionMode::OFF;
    float minMaterialForDiffuse = 4.0f;
    float minMaterialForSpecular = 4.0f;
    bool enableAntiFirefly = false;
    bool enableRoughnessEdgeStopping = true;
};

const uint32_t SIGMA_MAX_HISTORY_FRAME_NUM = 7;
const float SIGMA_DEFAULT_ACCUMULATION_TIME = 0.084f;

struct SigmaSettings {
    float lightDirection[3] = {0.0f, 0.0f, 0.0f};
    float planeDistanceSensitivity = 0.02f;
    uint32_t maxStabilizedFrameNum = 5;
};

const uint32_t REFERENCE_MAX_HISTORY_FRAME_NUM = 4095;
const float REFERENCE_DEFAULT_ACCUMULATION_TIME = 2.0f;

struct ReferenceSettings { uint32_t maxAccumulatedFrameNum = 120; };

// ---------------------------------------------------------------------------------------------
// Entry points (NRD.h:60-79). C++ references are pointers at the ABI level.
// ---------------------------------------------------------------------------------------------
NRD_API Result NRD_CALL CreateInstance(const InstanceCreationDesc& instanceCreationDesc, Instance*& instance);
NRD_API void NRD_CALL DestroyInstance(Instance& instance);
NRD_API const LibraryDesc* NRD_CALL GetLibraryDesc();
NRD_API const InstanceDesc* NRD_CALL GetInstanceDesc(const Instance& instance);
NRD_API Result NRD_CALL SetCommonSettings(Instance& instance, const CommonSettings& commonSettings);
NRD_API Result NRD_CALL SetDenoiserSettings(Instance& instance, Identifier identifier, const void* denoiserSettings);
NRD_API Result NRD_CALL GetComputeDispatches(Instance& instance, const Identifier* identifiers, uint32_t identifiersNum,
                                             const DispatchDesc*& dispatchDescs, uint32_t& dispatchDescsNum);
NRD_API const char* GetResourceTypeString(ResourceType resourceType);
NRD_API const char* GetDenoiserString(Denoiser denoiser);

// Sizes measured from the reference build (g++ 13.3, x86-64): tests/test_abi.py re-checks them against oracle/_ref.
static_assert(sizeof(CommonSettings) == 432, "CommonSettings layout drifted from NRDSettings.h:84-194");
static_assert(sizeof(ReblurSettings) == 120, "ReblurSettings layout drifted from NRDSettings.h:256-339");
static_assert(sizeof(RelaxSettings) == 148, "RelaxSettings layout drifted from NRDSettings.h:361-452");
static_assert(sizeof(SigmaSettings) == 20, "SigmaSettings layout drifted from NRDSettings.h:461-474");
static_assert(sizeof(DispatchDesc) == 56 && sizeof(PipelineDesc) == 320 && sizeof(ResourceDesc) == 12 && sizeof(InstanceDesc) == 112 &&
                  sizeof(LibraryDesc) == 40 && sizeof(InstanceCreationDesc) == 48,
              "descriptor layout drifted");

}  // namespace nrd
