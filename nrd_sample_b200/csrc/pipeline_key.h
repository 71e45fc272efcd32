// Which kernel a pipeline of the instance is, decided ONCE. The reference picks its pipelines when the integration is created
// ( External/NRD/Integration/NRDIntegration.hpp:199-239: one pipeline object per InstanceDesc::pipelines entry, dispatches carry pipelineIndex ); here the
// text of PipelineDesc::shaderIdentifier ( "<file>|KEY=VALUE|..." , InstanceImpl.h:59-67 ) is parsed into this POD by nrdcuCreate for every pipeline of
// the instance, and a dispatch only switches on the enums. The context-less nrdcuDispatch entry resolves per call.
#pragma once
#include <cstdint>

namespace nrdk {

enum PipelineFamily : uint8_t { FAMILY_UNKNOWN = 0, FAMILY_CLEAR, FAMILY_REBLUR, FAMILY_RELAX, FAMILY_SIGMA, FAMILY_REFERENCE };

enum PipelinePass : uint8_t {
    PASS_NONE = 0,
    // REBLUR ( Shaders.cfg:3-38 )
    REBLUR_CLASSIFY_TILES, REBLUR_HITDIST_RECONSTRUCTION, REBLUR_PREPASS, REBLUR_TEMPORAL_ACCUMULATION, REBLUR_HISTORY_FIX, REBLUR_BLUR, REBLUR_POST_BLUR,
    REBLUR_TEMPORAL_STABILIZATION, REBLUR_SPLIT_SCREEN, REBLUR_VALIDATION,
    // RELAX ( :40-61 )
    RELAX_CLASSIFY_TILES, RELAX_HITDIST_RECONSTRUCTION, RELAX_PREPASS, RELAX_TEMPORAL_ACCUMULATION, RELAX_HISTORY_FIX, RELAX_HISTORY_CLAMPING, RELAX_COPY,
    RELAX_ANTI_FIREFLY, RELAX_ATROUS_SMEM, RELAX_ATROUS, RELAX_SPLIT_SCREEN, RELAX_VALIDATION,
    // SIGMA ( :63-70 )
    SIGMA_CLASSIFY_TILES, SIGMA_SMOOTH_TILES, SIGMA_COPY, SIGMA_BLUR, SIGMA_TEMPORAL_STABILIZATION, SIGMA_SPLIT_SCREEN,
    // REFERENCE
    REFERENCE_TEMPORAL_ACCUMULATION, REFERENCE_COPY,
};

struct PipelineKey {
    PipelineFamily family = FAMILY_UNKNOWN;
    PipelinePass pass = PASS_NONE;
    uint8_t signal = 0;     // NRD_SIGNAL: 1 DIFF, 2 SPEC, 3 BOTH; 0 = the permutation has none
    uint8_t mode = 0;       // NRD_MODE: 0 RADIANCE, 1 SH, 2 OCCLUSION, 3 DO ( MODE_* of kernels/reblur_common.cuh )
    bool hasMode = false;
    bool mode5x5 = false;                  // MODE_5X5=1
    bool temporalStabilization = false;    // TEMPORAL_STABILIZATION=1
    bool firstPass = false;                // FIRST_PASS=1
    bool translucency = false;             // TRANSLUCENCY=1
    const char* id = "";                   // the identifier it was resolved from ( owned by the instance / the caller ): messages only
};

// FAMILY_UNKNOWN when the file or one of its defines is not a permutation this library has a kernel for
PipelineKey resolvePipeline(const char* shaderIdentifier);

}  // namespace nrdk
