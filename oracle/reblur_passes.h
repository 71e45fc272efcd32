// TEST INFRASTRUCTURE — CPU oracle (see hlsl_like.h). Binding tables and entry points of reblur_passes.cpp.
// Member order == shader register order == DispatchDesc::resources order (REBLUR_*.resources.hlsli).
// `signal` = NRD_SIGNAL of the permutation (1 DIFF, 2 SPEC, 3 BOTH): bindings of the lobe a denoiser does not have are never touched.
#pragma once
#include "reblur_shared.h"

namespace orc {

struct TaTextures {
    const Tex *gIn_Tiles, *gIn_Normal_Roughness, *gIn_ViewZ, *gIn_Mv, *gPrev_ViewZ, *gPrev_Normal_Roughness, *gPrev_InternalData, *gIn_DisocclusionThresholdMix,
        *gIn_DiffConfidence, *gIn_SpecConfidence, *gIn_Diff, *gIn_Spec, *gHistory_Diff, *gHistory_Spec, *gHistory_DiffFast, *gHistory_SpecFast,
        *gPrev_SpecHitDistForTracking, *gIn_SpecHitDistForTracking;
    Tex *gOut_Data1, *gOut_Diff, *gOut_Spec, *gOut_DiffFast, *gOut_SpecFast, *gOut_SpecHitDistForTracking, *gOut_Data2;
};

struct HfTextures {
    const Tex *gIn_Tiles, *gIn_Normal_Roughness, *gIn_Data1, *gIn_ViewZ, *gIn_Diff, *gIn_Spec, *gIn_DiffFast, *gIn_SpecFast, *gIn_SpecHitDistForTracking;
    Tex *gOut_Diff, *gOut_Spec, *gOut_DiffFast, *gOut_SpecFast;
};

struct TsTextures {
    const Tex *gIn_Tiles, *gIn_Normal_Roughness, *gIn_ViewZ, *gIn_Data1, *gIn_Data2, *gIn_SpecHitDistForTracking, *gIn_Diff, *gIn_Spec, *gHistory_DiffLumaStabilized,
        *gHistory_SpecLumaStabilized;
    Tex *gInOut_Mv, *gOut_InternalData, *gOut_Diff, *gOut_Spec, *gOut_DiffLumaStabilized, *gOut_SpecLumaStabilized;
};

void reblurClassifyTiles(const ReblurCB& cb, const Tex& gIn_ViewZ, Tex& gOut_Tiles, int gridW, int gridH);
void reblurHitDistReconstruction(const ReblurCB& cb, const Tex& gIn_Tiles, const Tex& gIn_Normal_Roughness, const Tex& gIn_ViewZ, const Tex& gIn_Diff, const Tex& gIn_Spec,
                                 Tex& gOut_Diff, Tex& gOut_Spec, int gridW, int gridH, int border, int signal = 3);
void reblurPrePass(const ReblurCB& cb, const Tex& gIn_Tiles, const Tex& gIn_Normal_Roughness, const Tex& gIn_ViewZ, const Tex& gIn_Diff, const Tex& gIn_Spec,
                   Tex& gOut_Diff, Tex& gOut_Spec, Tex& gOut_SpecHitDistForTracking, int gridW, int gridH, bool robustMirrorTest, int signal = 3);
void reblurBlur(const ReblurCB& cb, const Tex& gIn_Tiles, const Tex& gIn_Normal_Roughness, const Tex& gIn_ViewZ, const Tex& gIn_Data1, const Tex& gIn_Diff,
                const Tex& gIn_Spec, Tex& gOut_ViewZ, Tex& gOut_Diff, Tex& gOut_Spec, int gridW, int gridH, bool quads, bool robustMirrorTest, int signal = 3);
void reblurPostBlur(const ReblurCB& cb, const Tex& gIn_Tiles, const Tex& gIn_Normal_Roughness, const Tex& gIn_Data1, const Tex& gIn_ViewZ, const Tex& gIn_Diff,
                    const Tex& gIn_Spec, Tex& gOut_Normal_Roughness, Tex& gOut_Diff, Tex& gOut_Spec, Tex* gOut_InternalData, Tex* gOut_DiffCopy, Tex* gOut_SpecCopy,
                    bool temporalStabilization, int gridW, int gridH, bool quads, bool robustMirrorTest, int signal = 3);
void reblurTemporalAccumulation(const ReblurCB& cb, const TaTextures& t, int gridW, int gridH, int signal = 3);
void reblurHistoryFix(const ReblurCB& cb, const HfTextures& t, int gridW, int gridH, bool quads, int signal = 3);
void reblurTemporalStabilization(const ReblurCB& cb, const TsTextures& t, int gridW, int gridH, int signal = 3);
void clearTexture(Tex& out);

}  // namespace orc
