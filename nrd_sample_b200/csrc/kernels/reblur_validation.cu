// REBLUR validation overlay ( CommonSettings::enableValidation; replaces External/NRD/Shaders/REBLUR_Validation.cs.hlsl:36-340 ): a 4 x 4 grid of viewports over
// OUT_VALIDATION showing normals, roughness, viewZ, motion-vector error, world units + jitter + blur rotators, virtual history amount, accumulated frames per lobe
// and the input hit distances. One thread per OUT_VALIDATION texel; every input is point-sampled at the viewport's uv, so the pass is pure bandwidth
// ( 16 small pictures of the frame ) and not part of the timed denoising chain. Viewports this denoiser has nothing to show in keep what the texture held.
#include "debug_overlay.cuh"

namespace nrdk {

namespace {
__device__ constexpr float kSpecial8X[8] = {-1.0f, 0.0f, 1.0f, 0.0f, -0.35355339f, 0.35355339f, 0.35355339f, -0.35355339f};   // g_Special8 ( Common.hlsli:207-218 ): .xy
__device__ constexpr float kSpecial8Y[8] = {0.0f, 1.0f, 0.0f, -1.0f, 0.35355339f, 0.35355339f, -0.35355339f, -0.35355339f};

// first channel(s) of a texture of any bound format at the NEAREST texel of uv ( gNearestClamp )
// ( anyFetch4 plus the two-channel formats only this pass meets: data1 of a two-lobe denoiser is RG8_UNORM, an application may own RG16F motion vectors )
NRD_DEV float4 nearestAny(const TexView& t, float2 uv) {
    const int x = t.cx((int)floorf(uv.x * (float)t.w)), y = t.cy((int)floorf(uv.y * (float)t.h));
    const size_t i = (size_t)(y * t.pitch + x);
    switch ((nrd::Format)t.fmt) {
        case nrd::Format::RG8_UNORM: {
            const uchar2 v = __ldg(reinterpret_cast<const uchar2*>(t.data) + i);
            return make_float4((float)v.x / 255.0f, (float)v.y / 255.0f, 0.0f, 0.0f);
        }
        case nrd::Format::RG16_SFLOAT: {
            const float2 v = __half22float2(__ldg(reinterpret_cast<const __half2*>(t.data) + i));
            return make_float4(v.x, v.y, 0.0f, 0.0f);
        }
        case nrd::Format::RG32_SFLOAT: {
            const float2 v = __ldg(reinterpret_cast<const float2*>(t.data) + i);
            return make_float4(v.x, v.y, 0.0f, 0.0f);
        }
        case nrd::Format::R8_UINT: return make_float4((float)__ldg(t.data + i), 0.0f, 0.0f, 0.0f);
        case nrd::Format::R16_UINT: return make_float4((float)__ldg(reinterpret_cast<const unsigned short*>(t.data) + i), 0.0f, 0.0f, 0.0f);
        case nrd::Format::R32_UINT: return make_float4((float)__ldg(reinterpret_cast<const uint32_t*>(t.data) + i), 0.0f, 0.0f, 0.0f);
        default: return anyFetch4(t, x, y);
    }
}

__global__ void __launch_bounds__(256) reblurValidationKernel(const __grid_constant__ ReblurConstants cb, const __grid_constant__ ReblurValidationParams p) {
    pdlEntry();
    const int px = blockIdx.x * 32 + (threadIdx.x & 31), py = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (!p.out.inside(px, py)) return;
    if (cb.resetHistory != 0u) {
        anyStore4(p.out, px, py, f4(0.0f));
        return;
    }
    const bool hasDiffuse = cb._pad[0] != 0u, hasSpecular = cb._pad[1] != 0u;   // gHasDiffuse / gHasSpecular ride behind the shared constants ( Reblur.cpp:196-199 )
    const float2 resourceSize = make_float2(cb.resourceSize[0], cb.resourceSize[1]);
    const OverlayCell cell = overlayCell(px, py, resourceSize);
    const float2 uvScaled = cell.uv * make_float2(cb.resolutionScale[0], cb.resolutionScale[1]);

    const float4 nr = unpackNormalRoughness(p.normalRoughness.sampleNearestRaw(uvScaled));
    const float viewZ = unpackViewZ(cb, p.viewZ.sampleNearest(uvScaled));
    const float viewZraw = viewZ;   // the reference tests the sign AFTER UnpackViewZ ( an absolute value ): its "negative Z" caption and colour never show
    const float4 mvRaw = nearestAny(p.mv, uvScaled);
    const float3 mv = make_float3(mvRaw.x * cb.mvScale[0], mvRaw.y * cb.mvScale[1], mvRaw.z * cb.mvScale[2]);
    const float4 diff = nearestAny(p.diff, uvScaled * make_float2(cb.diffCheckerboard != 2u ? 0.5f : 1.0f, 1.0f));
    const float4 spec = nearestAny(p.spec, uvScaled * make_float2(cb.specCheckerboard != 2u ? 0.5f : 1.0f, 1.0f));
    // hit distance: .w of the lobe's input. The occlusion denoisers bind single-channel textures here, whose .w reads as 1 in HLSL: their two hit-distance
    // viewports are plain white wherever there is geometry. Kept as the reference has it.
    auto hitDistOf = [](const TexView& t, float4 v) {
        const nrd::Format f = (nrd::Format)t.fmt;
        return (f == nrd::Format::R8_UNORM || f == nrd::Format::R16_UNORM || f == nrd::Format::R16_SFLOAT || f == nrd::Format::R32_SFLOAT) ? 1.0f : v.w;
    };
    float4 d1 = nearestAny(p.data1, uvScaled);                                // RG8 ( two lobes ) or R8
    if (!hasDiffuse) d1.y = d1.x;
    const float2 data1 = make_float2(d1.x, d1.y) * REBLUR_MAX_ACCUM_FRAME_NUM;
    float virtualHistoryAmount = 0.0f;                                        // UnpackData2 with NRD_SPEC = 1 ( :19-20 ): 7 bits from bit 8
    {
        const int dx = (int)(uvScaled.x * resourceSize.x), dy = (int)(uvScaled.y * resourceSize.y);
        if (p.data2.inside(dx, dy)) {
            uint32_t bits = 0u;
            if (p.data2.fmt == (uint32_t)nrd::Format::R32_UINT) bits = __ldg(p.data2.ptr<uint32_t>(dx, dy));
            else if (p.data2.fmt == (uint32_t)nrd::Format::R8_UINT) bits = __ldg(p.data2.ptr<uint8_t>(dx, dy));
            virtualHistoryAmount = (float)((bits >> 8) & 127u) / 127.0f;
        }
    }
    const float3 N = xyz(nr);
    const float3 Xv = reconstructViewPosition(cell.uv, cb.frustum, viewZ, cb.orthoMode);
    const float3 X = rotate(cb.viewToWorld, Xv);
    const bool isInf = !inDenoisingRange(cb, viewZ);
    const bool checker = ((((unsigned)px >> 2) ^ ((unsigned)py >> 2)) & 1u) != 0u;   // Sequence::CheckerBoard( pixelPos >> 2, 0 )
    const float notInf = isInf ? 0.0f : 1.0f;

    Caption text(px, py, cell.captionX, cell.captionY);
    float4 result = anyFetch4(p.out, px, py);
    auto set = [&](float3 c) { result = f4(c, 1.0f); };

    switch (cell.index) {
        case 0:
            text.print("NORMALS-");
            text.nextChar();
            text.printUint(2u);   // NRD_NORMAL_ENCODING: R10G10B10A2_UNORM
            set(N * 0.5f + 0.5f);
            break;
        case 1:
            text.print("ROUGHNESS-");
            text.nextChar();
            text.printUint(1u);   // NRD_ROUGHNESS_ENCODING: linear
            set(f3(nr.w));
            break;
        case 2: {
            text.print("Z");
            if (viewZraw < 0.0f) text.print("-");
            const float f = 0.1f * viewZ / (1.0f + 0.1f * viewZ);
            set(isInf ? make_float3(1.0f, 0.0f, 0.0f) : (viewZraw < 0.0f ? make_float3(0.0f, 0.0f, f) : make_float3(0.0f, f, 0.0f)));
            break;
        }
        case 3: {
            text.print("MV");
            const float2 expected = screenUv(cb.worldToClipPrev, X);
            float2 prev = cell.uv + xy(mv);
            if (cb.mvScale[3] != 0.0f) prev = screenUv(cb.worldToClipPrev, X + mv);
            const float2 delta = (prev - expected) * make_float2(cb.rectSize[0], cb.rectSize[1]);
            set(isInScreenNearest(prev) ? make_float3(fabsf(delta.x), fabsf(delta.y), 0.0f) : make_float3(0.0f, 0.0f, 1.0f));
            break;
        }
        case 4: {
            text.print("UNITS");
            const float2 dim = make_float2(0.5f * resourceSize.y / resourceSize.x, 0.5f);
            const float2 dimInPixels = resourceSize * 0.25f * dim;
            const float2 remapped = (cell.uv - (1.0f - dim)) / dim, remapped2 = (cell.uv - make_float2(1.0f - dim.x, 0.0f)) / dim;
            float3 rgb = xyz(result);
            if (remapped.x > 0.0f && remapped.y > 0.0f) {
                const float2 uv = make_float2(cb.jitter[0], cb.jitter[1]) + 0.5f;
                const bool valid = saturate(uv.x) == uv.x && saturate(uv.y) == uv.y;
                const int ax = (int)(saturate(uv.x) * dimInPixels.x), ay = (int)(saturate(uv.y) * dimInPixels.y);
                const int bx = (int)(remapped.x * dimInPixels.x), by = (int)(remapped.y * dimInPixels.y);
                if (abs(ax - bx) <= 1 && abs(ay - by) <= 1 && valid) rgb = f3(0.66f);
                if (abs(ax - bx) <= 3 && abs(ay - by) <= 3 && !valid) rgb = make_float3(1.0f, 0.0f, 0.0f);
                result = f4(rgb, 1.0f);
            } else if (remapped2.x > 0.0f && remapped2.y > 0.0f) {
                // the tap pattern of blur ( red ) and post-blur ( green ) under this frame's rotators, at the area factor of the shown history length
                const int bx = (int)(remapped2.x * dimInPixels.x), by = (int)(remapped2.y * dimInPixels.y);
                const uint32_t maxFrames = (uint32_t)cb.maxAccumulatedFrameNum;
                const uint32_t frameIndex = maxFrames ? (cb.frameIndex >> 2) % maxFrames : 0u;
                const float scale = 0.5f / 2.0f * sqrt01(advancedNonLinearAccumSpeed(cb, (float)frameIndex));   // / max( blur, post-blur radius scale )
                float4 acc = result;
                for (int n = 0; n < 8; n++) {
                    const float2 o = make_float2(kSpecial8X[n], kSpecial8Y[n]) * scale;
                    const float2 u0 = rotate2(make_float4(cb.rotator[0], cb.rotator[1], cb.rotator[2], cb.rotator[3]), o * 1.0f) + 0.5f;
                    const float2 u1 = rotate2(make_float4(cb.rotatorPost[0], cb.rotatorPost[1], cb.rotatorPost[2], cb.rotatorPost[3]), o * 2.0f) + 0.5f;
                    if (abs((int)(saturate(u0.x) * dimInPixels.x) - bx) <= 1 && abs((int)(saturate(u0.y) * dimInPixels.y) - by) <= 1) acc.x += 1.0f;
                    if (abs((int)(saturate(u1.x) * dimInPixels.x) - bx) <= 1 && abs((int)(saturate(u1.y) * dimInPixels.y) - by) <= 1) acc.y += 1.0f;
                }
                result = frameIndex == 0u ? f4(0.0f) : saturate(acc);
                result.w = 1.0f;
            } else {
                const float3 shifted = X + viewZ * 0.001f;   // rounding error correction
                set(make_float3(frac(shifted.x), frac(shifted.y), frac(shifted.z)) * notInf);
            }
            break;
        }
        case 7:
            if (hasSpecular) {
                text.print("VIRTUAL HISTORY");
                set(f3(virtualHistoryAmount * notInf));
            }
            break;
        case 8:
        case 11:
            if (cell.index == 8 ? hasDiffuse : hasSpecular) {
                text.print(cell.index == 8 ? "DIFF FRAMES" : "SPEC FRAMES");
                const float frames = cell.index == 8 ? data1.x : data1.y;
                float f = 1.0f - saturate(frames / fmaxf(cb.maxAccumulatedFrameNum, 1.0f));
                if (checker && frames < 1.0f) f = 0.75f;
                set(colorizeZucconi(cell.uv.y > 0.95f ? 1.0f - cell.uv.x : f * notInf));
            }
            break;
        case 12:
        case 15:
            if (cell.index == 12 ? hasDiffuse : hasSpecular) {
                text.print(cell.index == 12 ? "DIFF HITT" : "SPEC HITT");
                const float h = cell.index == 12 ? hitDistOf(p.diff, diff) : hitDistOf(p.spec, spec);
                const float3 c = h == 0.0f ? make_float3(1.0f, 0.0f, 0.0f) : (h != saturate(h) ? make_float3(1.0f, 0.0f, 1.0f) : f3(h));
                set(c * notInf);
            }
            break;
        default: break;
    }
    float3 rgb = applyCaption(xyz(result), text.foreground);
    if (text.foreground && (cell.index == 12 || cell.index == 15)) rgb = f3(0.5f);   // hit distances are grey-scale: a flat caption reads better ( :335-336 )
    anyStore4(p.out, px, py, f4(rgb, result.w));
}
}  // namespace

void launchReblurValidation(const ReblurConstants& cb, const ReblurValidationParams& p, cudaStream_t stream) {
    launchK(reblurValidationKernel, dim3((p.out.w + 31) / 32, (p.out.h + 7) / 8), 256, 0, stream, cb, p);
}

}  // namespace nrdk
