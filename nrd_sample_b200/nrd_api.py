"""ctypes view of the nrd:: descriptor ABI (include/nrd_b200.h == External/NRD/Include/NRD*.h).

The same binding drives this repo's host library and — in tests only — the reference host library
built into oracle/_ref/libnrd_ref.so, because both export the nine NRD.h:60-79 symbols with identical
POD layouts. Plumbing only: no pixel math happens in Python.
"""
from __future__ import annotations

import ctypes as C
import enum
from dataclasses import dataclass
from typing import List, Optional, Sequence


class Result(enum.IntEnum):
    SUCCESS = 0
    FAILURE = 1
    INVALID_ARGUMENT = 2
    UNSUPPORTED = 3
    NON_UNIQUE_IDENTIFIER = 4


class ResourceType(enum.IntEnum):
    IN_MV = 0
    IN_NORMAL_ROUGHNESS = 1
    IN_VIEWZ = 2
    IN_DIFF_CONFIDENCE = 3
    IN_SPEC_CONFIDENCE = 4
    IN_DISOCCLUSION_THRESHOLD_MIX = 5
    IN_DIFF_RADIANCE_HITDIST = 6
    IN_SPEC_RADIANCE_HITDIST = 7
    IN_DIFF_HITDIST = 8
    IN_SPEC_HITDIST = 9
    IN_DIFF_DIRECTION_HITDIST = 10
    IN_DIFF_SH0 = 11
    IN_DIFF_SH1 = 12
    IN_SPEC_SH0 = 13
    IN_SPEC_SH1 = 14
    IN_PENUMBRA = 15
    IN_TRANSLUCENCY = 16
    IN_SIGNAL = 17
    OUT_DIFF_RADIANCE_HITDIST = 18
    OUT_SPEC_RADIANCE_HITDIST = 19
    OUT_DIFF_SH0 = 20
    OUT_DIFF_SH1 = 21
    OUT_SPEC_SH0 = 22
    OUT_SPEC_SH1 = 23
    OUT_DIFF_HITDIST = 24
    OUT_SPEC_HITDIST = 25
    OUT_DIFF_DIRECTION_HITDIST = 26
    OUT_SHADOW_TRANSLUCENCY = 27
    OUT_SIGNAL = 28
    OUT_VALIDATION = 29
    TRANSIENT_POOL = 30
    PERMANENT_POOL = 31


class Denoiser(enum.IntEnum):
    REBLUR_DIFFUSE = 0
    REBLUR_DIFFUSE_OCCLUSION = 1
    REBLUR_DIFFUSE_SH = 2
    REBLUR_SPECULAR = 3
    REBLUR_SPECULAR_OCCLUSION = 4
    REBLUR_SPECULAR_SH = 5
    REBLUR_DIFFUSE_SPECULAR = 6
    REBLUR_DIFFUSE_SPECULAR_OCCLUSION = 7
    REBLUR_DIFFUSE_SPECULAR_SH = 8
    REBLUR_DIFFUSE_DIRECTIONAL_OCCLUSION = 9
    RELAX_DIFFUSE = 10
    RELAX_DIFFUSE_SH = 11
    RELAX_SPECULAR = 12
    RELAX_SPECULAR_SH = 13
    RELAX_DIFFUSE_SPECULAR = 14
    RELAX_DIFFUSE_SPECULAR_SH = 15
    SIGMA_SHADOW = 16
    SIGMA_SHADOW_TRANSLUCENCY = 17
    REFERENCE = 18


class Format(enum.IntEnum):
    R8_UNORM = 0
    R8_UINT = 2
    RG8_UNORM = 4
    RGBA8_UNORM = 8
    R16_UNORM = 13
    R16_UINT = 15
    R16_SFLOAT = 17
    RG16_SFLOAT = 22
    RGBA16_SNORM = 24
    RGBA16_SFLOAT = 27
    R32_UINT = 28
    R32_SFLOAT = 30
    RG32_SFLOAT = 33
    RGBA32_SFLOAT = 39
    R10_G10_B10_A2_UNORM = 40


FORMAT_BYTES = {
    Format.R8_UNORM: 1, Format.R8_UINT: 1, Format.RG8_UNORM: 2, Format.RGBA8_UNORM: 4, Format.R16_UINT: 2, Format.R16_SFLOAT: 2,
    Format.RG16_SFLOAT: 4, Format.RGBA16_SFLOAT: 8, Format.R16_UNORM: 2, Format.RGBA16_SNORM: 8, Format.R32_UINT: 4, Format.R32_SFLOAT: 4, Format.R10_G10_B10_A2_UNORM: 4, Format.RGBA32_SFLOAT: 16,
}


class AccumulationMode(enum.IntEnum):
    CONTINUE = 0
    RESTART = 1
    CLEAR_AND_RESTART = 2


class DescriptorType(enum.IntEnum):
    TEXTURE = 0
    STORAGE_TEXTURE = 1


# ------------------------------------------------------------------------------------------------
# PODs
# ------------------------------------------------------------------------------------------------
class AllocationCallbacks(C.Structure):
    _fields_ = [("Allocate", C.c_void_p), ("Reallocate", C.c_void_p), ("Free", C.c_void_p), ("userArg", C.c_void_p)]


class DenoiserDesc(C.Structure):
    _fields_ = [("identifier", C.c_uint32), ("denoiser", C.c_uint32)]


class InstanceCreationDesc(C.Structure):
    _fields_ = [("allocationCallbacks", AllocationCallbacks), ("denoisers", C.POINTER(DenoiserDesc)), ("denoisersNum", C.c_uint32)]


class SPIRVBindingOffsets(C.Structure):
    _fields_ = [("samplerOffset", C.c_uint32), ("textureOffset", C.c_uint32), ("constantBufferOffset", C.c_uint32),
                ("storageTextureAndBufferOffset", C.c_uint32)]


class LibraryDesc(C.Structure):
    _fields_ = [("spirvBindingOffsets", SPIRVBindingOffsets), ("supportedDenoisers", C.POINTER(C.c_uint32)),
                ("supportedDenoisersNum", C.c_uint32), ("versionMajor", C.c_uint8), ("versionMinor", C.c_uint8),
                ("versionBuild", C.c_uint8), ("normalEncoding", C.c_uint8), ("roughnessEncoding", C.c_uint8)]


class TextureDesc(C.Structure):
    _fields_ = [("format", C.c_uint32), ("downsampleFactor", C.c_uint16)]


class ResourceDesc(C.Structure):
    _fields_ = [("descriptorType", C.c_uint32), ("type", C.c_uint32), ("indexInPool", C.c_uint16)]


class ResourceRangeDesc(C.Structure):
    _fields_ = [("descriptorType", C.c_uint32), ("descriptorsNum", C.c_uint32)]


class ComputeShaderDesc(C.Structure):
    _fields_ = [("bytecode", C.c_void_p), ("size", C.c_uint64)]


class PipelineDesc(C.Structure):
    _fields_ = [("computeShaderDXBC", ComputeShaderDesc), ("computeShaderDXIL", ComputeShaderDesc),
                ("computeShaderSPIRV", ComputeShaderDesc), ("resourceRanges", C.POINTER(ResourceRangeDesc)),
                ("resourceRangesNum", C.c_uint32), ("hasConstantData", C.c_bool), ("shaderIdentifier", C.c_char * 256)]


class DescriptorPoolDesc(C.Structure):
    _fields_ = [("perSetTexturesMaxNum", C.c_uint32), ("perSetStorageTexturesMaxNum", C.c_uint32), ("totalTexturesNum", C.c_uint32),
                ("totalStorageTexturesNum", C.c_uint32), ("setsMaxNum", C.c_uint32)]


class InstanceDesc(C.Structure):
    _fields_ = [("constantBufferAndSamplersSpaceIndex", C.c_uint32), ("resourcesSpaceIndex", C.c_uint32),
                ("constantBufferRegisterIndex", C.c_uint32), ("samplersBaseRegisterIndex", C.c_uint32),
                ("resourcesBaseRegisterIndex", C.c_uint32), ("constantBufferMaxDataSize", C.c_uint32),
                ("samplers", C.POINTER(C.c_uint32)), ("samplersNum", C.c_uint32), ("shaderEntryPoint", C.c_char_p),
                ("pipelines", C.POINTER(PipelineDesc)), ("pipelinesNum", C.c_uint32),
                ("permanentPool", C.POINTER(TextureDesc)), ("permanentPoolSize", C.c_uint32),
                ("transientPool", C.POINTER(TextureDesc)), ("transientPoolSize", C.c_uint32),
                ("descriptorPoolDesc", DescriptorPoolDesc)]


class DispatchDesc(C.Structure):
    _fields_ = [("name", C.c_char_p), ("identifier", C.c_uint32), ("resources", C.POINTER(ResourceDesc)),
                ("resourcesNum", C.c_uint32), ("constantBufferData", C.POINTER(C.c_uint8)), ("constantBufferDataSize", C.c_uint32),
                ("constantBufferDataMatchesPreviousDispatch", C.c_bool), ("pipelineIndex", C.c_uint16),
                ("gridWidth", C.c_uint16), ("gridHeight", C.c_uint16)]


_IDENTITY = (1.0, 0, 0, 0, 0, 1.0, 0, 0, 0, 0, 1.0, 0, 0, 0, 0, 1.0)


class CommonSettings(C.Structure):
    _fields_ = [("viewToClipMatrix", C.c_float * 16), ("viewToClipMatrixPrev", C.c_float * 16), ("worldToViewMatrix", C.c_float * 16),
                ("worldToViewMatrixPrev", C.c_float * 16), ("worldPrevToWorldMatrix", C.c_float * 16),
                ("motionVectorScale", C.c_float * 3), ("cameraJitter", C.c_float * 2), ("cameraJitterPrev", C.c_float * 2),
                ("resourceSize", C.c_uint16 * 2), ("resourceSizePrev", C.c_uint16 * 2), ("rectSize", C.c_uint16 * 2),
                ("rectSizePrev", C.c_uint16 * 2), ("viewZScale", C.c_float), ("timeDeltaBetweenFrames", C.c_float),
                ("denoisingRange", C.c_float), ("disocclusionThreshold", C.c_float), ("disocclusionThresholdAlternate", C.c_float),
                ("cameraAttachedReflectionMaterialID", C.c_float), ("strandMaterialID", C.c_float),
                ("historyFixAlternatePixelStrideMaterialID", C.c_float), ("strandThickness", C.c_float), ("splitScreen", C.c_float),
                ("printfAt", C.c_uint16 * 2), ("debug", C.c_float), ("rectOrigin", C.c_uint32 * 2), ("frameIndex", C.c_uint32),
                ("accumulationMode", C.c_uint8), ("isMotionVectorInWorldSpace", C.c_bool), ("isHistoryConfidenceAvailable", C.c_bool),
                ("isDisocclusionThresholdMixAvailable", C.c_bool), ("enableValidation", C.c_bool)]

    def __init__(self, **kw):
        super().__init__()
        self.worldPrevToWorldMatrix = (C.c_float * 16)(*_IDENTITY)
        self.motionVectorScale = (C.c_float * 3)(1.0, 1.0, 0.0)
        self.viewZScale = 1.0
        self.denoisingRange = 500000.0
        self.disocclusionThreshold = 0.01
        self.disocclusionThresholdAlternate = 0.05
        self.cameraAttachedReflectionMaterialID = 999.0
        self.strandMaterialID = 999.0
        self.historyFixAlternatePixelStrideMaterialID = 999.0
        self.strandThickness = 80e-6
        self.printfAt = (C.c_uint16 * 2)(9999, 9999)
        for k, v in kw.items():
            setattr(self, k, v)


class ReblurSettings(C.Structure):
    _fields_ = [("hitDistanceParameters", C.c_float * 3), ("antilagLuminanceSigmaScale", C.c_float), ("antilagLuminanceSensitivity", C.c_float),
                ("responsiveRoughnessThreshold", C.c_float), ("responsiveMinAccumulatedFrameNum", C.c_uint32),
                ("convergenceS", C.c_float), ("convergenceB", C.c_float), ("convergenceP", C.c_float),
                ("maxAccumulatedFrameNum", C.c_uint32), ("maxFastAccumulatedFrameNum", C.c_uint32), ("maxStabilizedFrameNum", C.c_uint32),
                ("historyFixFrameNum", C.c_uint32), ("historyFixBasePixelStride", C.c_uint32), ("historyFixAlternatePixelStride", C.c_uint32),
                ("fastHistoryClampingSigmaScale", C.c_float), ("diffusePrepassBlurRadius", C.c_float), ("specularPrepassBlurRadius", C.c_float),
                ("minHitDistanceWeight", C.c_float), ("minBlurRadius", C.c_float), ("maxBlurRadius", C.c_float),
                ("lobeAngleFraction", C.c_float), ("roughnessFraction", C.c_float), ("planeDistanceSensitivity", C.c_float),
                ("fireflySuppressorMinRelativeScale", C.c_float), ("minMaterialForDiffuse", C.c_float), ("minMaterialForSpecular", C.c_float),
                ("checkerboardMode", C.c_uint8), ("hitDistanceReconstructionMode", C.c_uint8), ("enableAntiFirefly", C.c_bool),
                ("usePrepassOnlyForSpecularMotionEstimation", C.c_bool), ("returnHistoryLengthInsteadOfOcclusion", C.c_bool)]

    def __init__(self, **kw):
        super().__init__()
        self.hitDistanceParameters = (C.c_float * 3)(3.0, 0.1, 20.0)
        self.antilagLuminanceSigmaScale = 2.0
        self.antilagLuminanceSensitivity = 3.0
        self.responsiveRoughnessThreshold = 0.0
        self.responsiveMinAccumulatedFrameNum = 3
        self.convergenceS, self.convergenceB, self.convergenceP = 1.0, 0.2, 0.8
        self.maxAccumulatedFrameNum = 30
        self.maxFastAccumulatedFrameNum = 6
        self.maxStabilizedFrameNum = 63
        self.historyFixFrameNum = 3
        self.historyFixBasePixelStride = 14
        self.historyFixAlternatePixelStride = 14
        self.fastHistoryClampingSigmaScale = 2.0
        self.diffusePrepassBlurRadius = 30.0
        self.specularPrepassBlurRadius = 50.0
        self.minHitDistanceWeight = 0.1
        self.minBlurRadius = 1.0
        self.maxBlurRadius = 30.0
        self.lobeAngleFraction = 0.15
        self.roughnessFraction = 0.15
        self.planeDistanceSensitivity = 0.02
        self.fireflySuppressorMinRelativeScale = 2.0
        self.minMaterialForDiffuse = 4.0
        self.minMaterialForSpecular = 4.0
        self.enableAntiFirefly = True
        for k, v in kw.items():
            setattr(self, k, v)


class RelaxSettings(C.Structure):
    """nrd::RelaxSettings (NRDSettings.h:361-452), 148 bytes."""
    _fields_ = [("antilagAccelerationAmount", C.c_float), ("antilagSpatialSigmaScale", C.c_float), ("antilagTemporalSigmaScale", C.c_float), ("antilagResetAmount", C.c_float),
                ("diffuseMaxAccumulatedFrameNum", C.c_uint32), ("specularMaxAccumulatedFrameNum", C.c_uint32),
                ("diffuseMaxFastAccumulatedFrameNum", C.c_uint32), ("specularMaxFastAccumulatedFrameNum", C.c_uint32),
                ("historyFixFrameNum", C.c_uint32), ("historyFixBasePixelStride", C.c_uint32), ("historyFixAlternatePixelStride", C.c_uint32),
                ("historyFixEdgeStoppingNormalPower", C.c_float), ("fastHistoryClampingSigmaScale", C.c_float),
                ("diffusePrepassBlurRadius", C.c_float), ("specularPrepassBlurRadius", C.c_float), ("minHitDistanceWeight", C.c_float),
                ("spatialVarianceEstimationHistoryThreshold", C.c_uint32), ("diffusePhiLuminance", C.c_float), ("specularPhiLuminance", C.c_float),
                ("lobeAngleFraction", C.c_float), ("roughnessFraction", C.c_float), ("specularVarianceBoost", C.c_float), ("specularLobeAngleSlack", C.c_float),
                ("atrousIterationNum", C.c_uint32), ("diffuseMinLuminanceWeight", C.c_float), ("specularMinLuminanceWeight", C.c_float), ("depthThreshold", C.c_float),
                ("confidenceDrivenRelaxationMultiplier", C.c_float), ("confidenceDrivenLuminanceEdgeStoppingRelaxation", C.c_float),
                ("confidenceDrivenNormalEdgeStoppingRelaxation", C.c_float), ("luminanceEdgeStoppingRelaxation", C.c_float),
                ("normalEdgeStoppingRelaxation", C.c_float), ("roughnessEdgeStoppingRelaxation", C.c_float),
                ("checkerboardMode", C.c_uint8), ("hitDistanceReconstructionMode", C.c_uint8),
                ("minMaterialForDiffuse", C.c_float), ("minMaterialForSpecular", C.c_float), ("enableAntiFirefly", C.c_bool), ("enableRoughnessEdgeStopping", C.c_bool)]

    def __init__(self, **kw):
        super().__init__()
        self.antilagAccelerationAmount, self.antilagSpatialSigmaScale, self.antilagTemporalSigmaScale, self.antilagResetAmount = 0.3, 4.5, 0.5, 0.5
        self.diffuseMaxAccumulatedFrameNum = self.specularMaxAccumulatedFrameNum = 30
        self.diffuseMaxFastAccumulatedFrameNum = self.specularMaxFastAccumulatedFrameNum = 6
        self.historyFixFrameNum = 3
        self.historyFixBasePixelStride = self.historyFixAlternatePixelStride = 14
        self.historyFixEdgeStoppingNormalPower = 8.0
        self.fastHistoryClampingSigmaScale = 2.0
        self.diffusePrepassBlurRadius, self.specularPrepassBlurRadius = 30.0, 50.0
        self.minHitDistanceWeight = 0.1
        self.spatialVarianceEstimationHistoryThreshold = 3
        self.diffusePhiLuminance, self.specularPhiLuminance = 2.0, 1.0
        self.lobeAngleFraction, self.roughnessFraction = 0.5, 0.15
        self.specularLobeAngleSlack = 0.15
        self.atrousIterationNum = 5
        self.depthThreshold = 0.003
        self.luminanceEdgeStoppingRelaxation, self.normalEdgeStoppingRelaxation, self.roughnessEdgeStoppingRelaxation = 0.5, 0.3, 1.0
        self.minMaterialForDiffuse = self.minMaterialForSpecular = 4.0
        self.enableRoughnessEdgeStopping = True
        for k, v in kw.items():
            setattr(self, k, v)


class ReferenceSettings(C.Structure):
    """nrd::ReferenceSettings (NRDSettings.h:483-487)."""
    _fields_ = [("maxAccumulatedFrameNum", C.c_uint32)]

    def __init__(self, **kw):
        super().__init__()
        self.maxAccumulatedFrameNum = 120
        for k, v in kw.items():
            setattr(self, k, v)


class SigmaSettings(C.Structure):
    _fields_ = [("lightDirection", C.c_float * 3), ("planeDistanceSensitivity", C.c_float), ("maxStabilizedFrameNum", C.c_uint32)]

    def __init__(self, **kw):
        super().__init__()
        self.planeDistanceSensitivity = 0.02
        self.maxStabilizedFrameNum = 5
        for k, v in kw.items():
            setattr(self, k, v)


assert C.sizeof(CommonSettings) == 432 and C.sizeof(ReblurSettings) == 120 and C.sizeof(SigmaSettings) == 20 and C.sizeof(RelaxSettings) == 148
assert C.sizeof(DispatchDesc) == 56 and C.sizeof(PipelineDesc) == 320 and C.sizeof(InstanceDesc) == 112

EXPORTED_SYMBOLS = ("CreateInstance", "DestroyInstance", "GetLibraryDesc", "GetInstanceDesc", "SetCommonSettings",
                    "SetDenoiserSettings", "GetComputeDispatches", "GetResourceTypeString", "GetDenoiserString")


# ------------------------------------------------------------------------------------------------
# Plain-python snapshots (descriptor memory is owned by the instance and overwritten per call)
# ------------------------------------------------------------------------------------------------
@dataclass
class Binding:
    descriptor: int   # DescriptorType
    type: int         # ResourceType
    index: int


@dataclass
class Dispatch:
    name: str
    identifier: int
    shader: str
    bindings: List[Binding]
    constants: bytes
    constants_match_previous: bool
    pipeline_index: int
    grid: tuple


class NrdLibrary:
    """One loaded host library exporting the NRD.h entry points."""

    def __init__(self, path: str):
        self.path = path
        self.lib = C.CDLL(path)
        L = self.lib
        L.CreateInstance.argtypes = [C.POINTER(InstanceCreationDesc), C.POINTER(C.c_void_p)]
        L.CreateInstance.restype = C.c_uint32
        L.DestroyInstance.argtypes = [C.c_void_p]
        L.DestroyInstance.restype = None
        L.GetLibraryDesc.restype = C.POINTER(LibraryDesc)
        L.GetInstanceDesc.argtypes = [C.c_void_p]
        L.GetInstanceDesc.restype = C.POINTER(InstanceDesc)
        L.SetCommonSettings.argtypes = [C.c_void_p, C.POINTER(CommonSettings)]
        L.SetCommonSettings.restype = C.c_uint32
        L.SetDenoiserSettings.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        L.SetDenoiserSettings.restype = C.c_uint32
        L.GetComputeDispatches.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.c_uint32, C.POINTER(C.POINTER(DispatchDesc)), C.POINTER(C.c_uint32)]
        L.GetComputeDispatches.restype = C.c_uint32
        L.GetResourceTypeString.argtypes = [C.c_uint32]
        L.GetResourceTypeString.restype = C.c_char_p
        L.GetDenoiserString.argtypes = [C.c_uint32]
        L.GetDenoiserString.restype = C.c_char_p

    def library_desc(self) -> LibraryDesc:
        return self.lib.GetLibraryDesc().contents

    def supported_denoisers(self) -> List[int]:
        d = self.library_desc()
        return [d.supportedDenoisers[i] for i in range(d.supportedDenoisersNum)]


class NrdInstance:
    """nrd::Instance wrapper: CreateInstance .. GetComputeDispatches."""

    def __init__(self, library: NrdLibrary, denoisers: Sequence[tuple]):
        self.library = library
        arr = (DenoiserDesc * len(denoisers))(*[DenoiserDesc(int(i), int(d)) for i, d in denoisers])
        desc = InstanceCreationDesc()
        desc.denoisers = arr
        desc.denoisersNum = len(denoisers)
        handle = C.c_void_p()
        self.result = Result(library.lib.CreateInstance(C.byref(desc), C.byref(handle)))
        self.handle: Optional[C.c_void_p] = handle if self.result == Result.SUCCESS else None
        self._keep = arr

    def destroy(self):
        if self.handle:
            self.library.lib.DestroyInstance(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    def desc(self) -> InstanceDesc:
        return self.library.lib.GetInstanceDesc(self.handle).contents

    def pools(self):
        d = self.desc()
        perm = [(d.permanentPool[i].format, d.permanentPool[i].downsampleFactor) for i in range(d.permanentPoolSize)]
        tran = [(d.transientPool[i].format, d.transientPool[i].downsampleFactor) for i in range(d.transientPoolSize)]
        return perm, tran

    def shader_identifiers(self) -> List[str]:
        d = self.desc()
        return [d.pipelines[i].shaderIdentifier.decode() for i in range(d.pipelinesNum)]

    def set_common_settings(self, cs: CommonSettings) -> Result:
        return Result(self.library.lib.SetCommonSettings(self.handle, C.byref(cs)))

    def set_denoiser_settings(self, identifier: int, settings: C.Structure) -> Result:
        return Result(self.library.lib.SetDenoiserSettings(self.handle, identifier, C.byref(settings)))

    def get_compute_dispatches(self, identifiers: Sequence[int]):
        ids = (C.c_uint32 * max(len(identifiers), 1))(*identifiers)
        out = C.POINTER(DispatchDesc)()
        n = C.c_uint32()
        r = Result(self.library.lib.GetComputeDispatches(self.handle, ids if identifiers else None, len(identifiers), C.byref(out), C.byref(n)))
        shaders = self.shader_identifiers() if n.value else []
        dispatches = []
        for i in range(n.value):
            dd = out[i]
            bindings = [Binding(dd.resources[j].descriptorType, dd.resources[j].type, dd.resources[j].indexInPool) for j in range(dd.resourcesNum)]
            cb = bytes(dd.constantBufferData[: dd.constantBufferDataSize]) if dd.constantBufferDataSize and dd.constantBufferData else b""
            dispatches.append(Dispatch(dd.name.decode(), dd.identifier, shaders[dd.pipelineIndex], bindings, cb,
                                       bool(dd.constantBufferDataMatchesPreviousDispatch), dd.pipelineIndex, (dd.gridWidth, dd.gridHeight)))
        return r, dispatches
