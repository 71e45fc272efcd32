// Compile-and-run check of include/NRDIntegrationCuda.h ( tests/test_abi.py ): the call sequence NRDSample uses around nrd::Integration
// ( Source/NRDSample.cpp: Recreate at start-up, NewFrame / SetCommonSettings / SetDenoiserSettings / Denoise per frame, Destroy ), against the
// CUDA twin. Without a device Recreate must fail loudly; with one ( argv[1] == "gpu" ) two frames of REBLUR_DIFFUSE_SPECULAR are denoised.
#include <cuda_runtime_api.h>

#include <cstdio>
#include <vector>

#include "NRDIntegrationCuda.h"

int main(int argc, char** argv) {
    const bool gpu = argc > 1 && !strcmp(argv[1], "gpu");
    const uint16_t W = 256, H = 144;
    const nrd::Identifier id = 7;
    const nrd::DenoiserDesc denoisers[] = {{id, nrd::Denoiser::REBLUR_DIFFUSE_SPECULAR}};
    nrd::InstanceCreationDesc instanceDesc = {};
    instanceDesc.denoisers = denoisers;
    instanceDesc.denoisersNum = 1;
    nrd::IntegrationCudaCreationDesc desc = {};
    snprintf(desc.name, sizeof(desc.name), "NRD");
    desc.resourceWidth = W;
    desc.resourceHeight = H;

    nrd::IntegrationCuda NRD;
    nrd::Result r = NRD.Recreate(desc, instanceDesc, 0);
    if (!gpu) {
        // no CUDA device here: the twin must say so instead of pretending ( there is no CPU fallback )
        printf("Recreate -> %u (%s)\n", (unsigned)r, NRD.GetLastError());
        return r == nrd::Result::SUCCESS ? 1 : 0;
    }
    if (r != nrd::Result::SUCCESS) { printf("Recreate failed: %s\n", NRD.GetLastError()); return 2; }
    printf("memory: total %.2f MB = persistent %.2f + aliasable %.2f ( + private %.2f )\n", NRD.GetTotalMemoryUsageInMb(), NRD.GetPersistentMemoryUsageInMb(),
           NRD.GetAliasableMemoryUsageInMb(), NRD.GetPrivateMemoryUsageInMb());

    struct Tex { nrd::ResourceType slot; nrd::Format format; uint32_t bpp; };
    const Tex textures[] = {{nrd::ResourceType::IN_MV, nrd::Format::RGBA16_SFLOAT, 8}, {nrd::ResourceType::IN_NORMAL_ROUGHNESS, nrd::Format::R10_G10_B10_A2_UNORM, 4},
                            {nrd::ResourceType::IN_VIEWZ, nrd::Format::R32_SFLOAT, 4}, {nrd::ResourceType::IN_DIFF_RADIANCE_HITDIST, nrd::Format::RGBA16_SFLOAT, 8},
                            {nrd::ResourceType::IN_SPEC_RADIANCE_HITDIST, nrd::Format::RGBA16_SFLOAT, 8}, {nrd::ResourceType::OUT_DIFF_RADIANCE_HITDIST, nrd::Format::RGBA16_SFLOAT, 8},
                            {nrd::ResourceType::OUT_SPEC_RADIANCE_HITDIST, nrd::Format::RGBA16_SFLOAT, 8}};
    nrd::ResourceSnapshotCuda snapshot;
    std::vector<void*> allocations;
    for (const Tex& t : textures) {
        void* p = nullptr;
        if (cudaMalloc(&p, (size_t)W * H * t.bpp) != cudaSuccess) return 3;
        if (t.slot == nrd::ResourceType::IN_VIEWZ) {
            std::vector<float> z((size_t)W * H, 5.0f);
            cudaMemcpy(p, z.data(), z.size() * 4, cudaMemcpyHostToDevice);
        } else
            cudaMemset(p, 0, (size_t)W * H * t.bpp);
        allocations.push_back(p);
        nrd::ResourceCuda res;
        res.data = p; res.width = W; res.height = H; res.pitchBytes = W * t.bpp; res.format = t.format;
        snapshot.SetResource(t.slot, res);
    }
    nrd::CommonSettings common = {};
    const float proj[16] = {1.0f, 0, 0, 0, 0, 1.7f, 0, 0, 0, 0, 0, 1.0f, 0, 0, 0.1f, 0};   // left-handed perspective, infinite far plane
    memcpy(common.viewToClipMatrix, proj, 64);
    memcpy(common.viewToClipMatrixPrev, proj, 64);
    common.resourceSize[0] = common.resourceSizePrev[0] = common.rectSize[0] = common.rectSizePrev[0] = W;
    common.resourceSize[1] = common.resourceSizePrev[1] = common.rectSize[1] = common.rectSizePrev[1] = H;
    common.motionVectorScale[0] = 1.0f / W;
    common.motionVectorScale[1] = 1.0f / H;
    nrd::ReblurSettings reblur = {};
    for (uint32_t frame = 0; frame < 2; frame++) {
        NRD.NewFrame();
        common.frameIndex = frame;
        if (NRD.SetCommonSettings(common) != nrd::Result::SUCCESS) { printf("SetCommonSettings: %s\n", NRD.GetLastError()); return 4; }
        if (NRD.SetDenoiserSettings(id, &reblur) != nrd::Result::SUCCESS) { printf("SetDenoiserSettings: %s\n", NRD.GetLastError()); return 5; }
        if (NRD.Denoise(&id, 1, nullptr, snapshot) != nrd::Result::SUCCESS) { printf("Denoise: %s\n", NRD.GetLastError()); return 6; }
    }
    if (cudaDeviceSynchronize() != cudaSuccess) return 7;
    printf("denoised %u frames, %llu kernel launches\n", NRD.GetFrameIndex(), (unsigned long long)nrdcuGetLaunchCount());
    NRD.Destroy();
    for (void* p : allocations) cudaFree(p);
    return 0;
}
