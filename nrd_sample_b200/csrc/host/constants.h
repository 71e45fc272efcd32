// Constant-buffer layouts shared by the host pass graph (which fills them) and the CUDA kernels
// (which read them from __grid_constant__ parameters).
//
// Field order restates the reference's cbuffer macro lists so that the bytes a DispatchDesc carries
// are interchangeable with the reference library's:
//   REBLUR  — External/NRD/Shaders/REBLUR_Config.hlsli:115-192 (864 bytes)
//   SIGMA   — External/NRD/Shaders/SIGMA_Config.hlsli:44-78    (528 bytes)
//   RELAX   — External/NRD/Shaders/RELAX_Config.hlsli:21-101   (720 bytes; the a-trous passes append gStepSize, gIsLastPass)
// HLSL packing rules in force: float4x4 is column-major (64 B), float4 16 B, float2/uint2/int2 8 B,
// scalars 4 B, struct padded to 16 B.
#pragma once
#include <cstdint>

namespace nrdb {

struct Mat4 {
    float m[16];  // column-major: element(row r, col c) = m[c * 4 + r]
};

struct ReblurConstants {
    Mat4 worldToClip;
    Mat4 viewToClip;
    Mat4 viewToWorld;
    Mat4 worldToViewPrev;
    Mat4 worldToClipPrev;
    Mat4 worldPrevToWorld;
    float rotatorPre[4];
    float rotator[4];
    float rotatorPost[4];
    float frustum[4];
    float frustumPrev[4];
    float cameraDelta[4];
    float hitDistSettings[4];
    float viewVectorWorld[4];
    float viewVectorWorldPrev[4];
    float mvScale[4];
    float convergenceSettings[4];
    float antilagSettings[2];
    float resourceSize[2];
    float resourceSizeInv[2];
    float resourceSizeInvPrev[2];
    float rectSize[2];
    float rectSizeInv[2];
    float rectSizePrev[2];
    float resolutionScale[2];
    float resolutionScalePrev[2];
    float rectOffset[2];
    float jitter[2];
    uint32_t printfAt[2];
    uint32_t rectOrigin[2];
    int32_t rectSizeMinusOne[2];
    float disocclusionThreshold;
    float disocclusionThresholdAlternate;
    float cameraAttachedReflectionMaterialID;
    float strandMaterialID;
    float strandThickness;
    float stabilizationStrength;
    float debug;
    float orthoMode;
    float unproject;
    float denoisingRange;
    float planeDistSensitivity;
    float framerateScale;
    float minBlurRadius;
    float maxBlurRadius;
    float diffPrepassBlurRadius;
    float specPrepassBlurRadius;
    float maxAccumulatedFrameNum;
    float maxFastAccumulatedFrameNum;
    float antiFirefly;
    float lobeAngleFraction;
    float roughnessFraction;
    float historyFixFrameNum;
    float historyFixBasePixelStride;
    float historyFixAlternatePixelStride;
    float historyFixAlternatePixelStrideMaterialID;
    float fastHistoryClampingSigmaScale;
    float minRectDimMulUnproject;
    float usePrepassNotOnlyForSpecularMotionEstimation;
    float splitScreen;
    float splitScreenPrev;
    float checkerboardResolveAccumSpeed;
    float viewZScale;
    float fireflySuppressorMinRelativeScale;
    float minHitDistanceWeight;
    float diffMinMaterial;
    float specMinMaterial;
    float responsiveAccumulationInvRoughnessThreshold;
    uint32_t responsiveAccumulationMinAccumulatedFrameNum;
    uint32_t hasHistoryConfidence;
    uint32_t hasDisocclusionThresholdMix;
    uint32_t diffCheckerboard;
    uint32_t specCheckerboard;
    uint32_t frameIndex;
    uint32_t isRectChanged;
    uint32_t resetHistory;
    uint32_t returnHistoryLengthInsteadOfOcclusion;
    uint32_t _pad[2];
};
static_assert(sizeof(ReblurConstants) == 864, "REBLUR cbuffer must stay 864 bytes");

struct SigmaConstants {
    Mat4 worldToView;
    Mat4 viewToClip;
    Mat4 worldToClipPrev;
    Mat4 worldToViewPrev;
    float rotator[4];
    float rotatorPost[4];
    float viewVectorWorld[4];
    float lightDirectionView[4];
    float frustum[4];
    float frustumPrev[4];
    float cameraDelta[4];
    float mvScale[4];
    float resourceSizeInv[2];
    float resourceSizeInvPrev[2];
    float rectSize[2];
    float rectSizeInv[2];
    float rectSizePrev[2];
    float resolutionScale[2];
    float rectOffset[2];
    uint32_t printfAt[2];
    uint32_t rectOrigin[2];
    int32_t rectSizeMinusOne[2];
    int32_t tilesSizeMinusOne[2];
    float orthoMode;
    float unproject;
    float denoisingRange;
    float planeDistSensitivity;
    float stabilizationStrength;
    float debug;
    float splitScreen;
    float viewZScale;
    float minRectDimMulUnproject;
    uint32_t frameIndex;
    uint32_t isRectChanged;
    uint32_t _pad[3];
};
static_assert(sizeof(SigmaConstants) == 528, "SIGMA cbuffer must stay 528 bytes");

struct RelaxConstants {
    Mat4 worldToClip;
    Mat4 worldToClipPrev;
    Mat4 worldToViewPrev;
    Mat4 worldPrevToWorld;
    float rotatorPre[4];
    float frustumRight[4];
    float frustumUp[4];
    float frustumForward[4];
    float prevFrustumRight[4];
    float prevFrustumUp[4];
    float prevFrustumForward[4];
    float cameraDelta[4];
    float mvScale[4];
    float jitter[2];
    float resolutionScale[2];
    float rectOffset[2];
    float resourceSizeInv[2];
    float resourceSize[2];
    float rectSizeInv[2];
    float rectSizePrev[2];
    float resourceSizeInvPrev[2];
    uint32_t printfAt[2];
    uint32_t rectOrigin[2];
    int32_t rectSize[2];
    float specMaxAccumulatedFrameNum;
    float specMaxFastAccumulatedFrameNum;
    float diffMaxAccumulatedFrameNum;
    float diffMaxFastAccumulatedFrameNum;
    float disocclusionThreshold;
    float disocclusionThresholdAlternate;
    float cameraAttachedReflectionMaterialID;
    float strandMaterialID;
    float strandThickness;
    float roughnessFraction;
    float specVarianceBoost;
    float splitScreen;
    float diffBlurRadius;
    float specBlurRadius;
    float depthThreshold;
    float lobeAngleFraction;
    float specLobeAngleSlack;
    float historyFixEdgeStoppingNormalPower;
    float roughnessEdgeStoppingRelaxation;
    float normalEdgeStoppingRelaxation;
    float fastHistoryClampingSigmaScale;
    float historyAccelerationAmount;
    float historyResetTemporalSigmaScale;
    float historyResetSpatialSigmaScale;
    float historyResetAmount;
    float denoisingRange;
    float specPhiLuminance;
    float diffPhiLuminance;
    float diffMaxLuminanceRelativeDifference;
    float specMaxLuminanceRelativeDifference;
    float luminanceEdgeStoppingRelaxation;
    float confidenceDrivenRelaxationMultiplier;
    float confidenceDrivenLuminanceEdgeStoppingRelaxation;
    float confidenceDrivenNormalEdgeStoppingRelaxation;
    float debug;
    float orthoMode;
    float unproject;
    float framerateScale;
    float checkerboardResolveAccumSpeed;
    float historyFixFrameNum;
    float historyFixBasePixelStride;
    float historyFixAlternatePixelStride;
    float historyFixAlternatePixelStrideMaterialID;
    float historyThreshold;
    float viewZScale;
    float minHitDistanceWeight;
    float diffMinMaterial;
    float specMinMaterial;
    uint32_t roughnessEdgeStoppingEnabled;
    uint32_t frameIndex;
    uint32_t diffCheckerboard;
    uint32_t specCheckerboard;
    uint32_t hasHistoryConfidence;
    uint32_t hasDisocclusionThresholdMix;
    uint32_t resetHistory;
    uint32_t stepSize;    // RELAX_Atrous / RELAX_AtrousSmem only (RELAX_Atrous.resources.hlsli:11-15); padding otherwise
    uint32_t isLastPass;  // "
    uint32_t _pad[1];
};
static_assert(sizeof(RelaxConstants) == 720, "RELAX cbuffer must stay 720 bytes");

// REFERENCE_TemporalAccumulation.resources.hlsli:10-17, REFERENCE_Copy.resources.hlsli:10-18
struct ReferenceAccumulateConstants {
    float accumSpeed, debug, viewZScale, denoisingRange;
};
struct ReferenceCopyConstants {
    float rectSizeInv[2], splitScreen, debug, viewZScale, denoisingRange;
};
static_assert(sizeof(ReferenceAccumulateConstants) == 16 && sizeof(ReferenceCopyConstants) == 24, "REFERENCE cbuffers");

}  // namespace nrdb
