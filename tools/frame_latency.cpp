// Host-side cost of one frame where the kernels are tiny: SIGMA_SHADOW at 512 x 512 ( BASELINE.json configs[0] ) and REBLUR_DIFFUSE_SPECULAR at 256 x 144, driven
// from C++ through include/NRDIntegrationCuda.h exactly like an application would, with and without NRDCU_FLAG_CUDA_GRAPH. Prints one JSON line:
// microseconds per frame ( CUDA events around `frames` back-to-back frames, after a warm-up ) and the host time per Denoise call.
//   g++ -std=c++17 -O2 -I include -I $CUDA/include tools/frame_latency.cpp nrd_sample_b200/libnrd_b200.so -L $CUDA/lib64 -lcudart -Wl,-rpath,$PWD/nrd_sample_b200
#include <cuda_runtime_api.h>

#include <chrono>
#include <cstdio>
#include <vector>

#include "NRDIntegrationCuda.h"

struct Tex { nrd::ResourceType slot; nrd::Format format; uint32_t bpp; };

static double run(nrd::Denoiser denoiser, uint16_t W, uint16_t H, const std::vector<Tex>& textures, uint32_t flags, int frames, double* hostUs, uint64_t stats[3]) {
    const nrd::Identifier id = 1;
    const nrd::DenoiserDesc denoisers[] = {{id, denoiser}};
    nrd::InstanceCreationDesc instanceDesc = {};
    instanceDesc.denoisers = denoisers;
    instanceDesc.denoisersNum = 1;
    nrd::IntegrationCudaCreationDesc desc = {};
    snprintf(desc.name, sizeof(desc.name), "latency");
    desc.resourceWidth = W;
    desc.resourceHeight = H;
    desc.flags = flags;
    nrd::IntegrationCuda NRD;
    if (NRD.Recreate(desc, instanceDesc, 0) != nrd::Result::SUCCESS) { printf("Recreate failed: %s\n", NRD.GetLastError()); return -1.0; }
    nrd::ResourceSnapshotCuda snapshot;
    std::vector<void*> allocations;
    for (const Tex& t : textures) {
        void* p = nullptr;
        cudaMalloc(&p, (size_t)W * H * t.bpp);
        if (t.slot == nrd::ResourceType::IN_VIEWZ) {
            std::vector<float> z((size_t)W * H, 5.0f);
            cudaMemcpy(p, z.data(), z.size() * 4, cudaMemcpyHostToDevice);
        } else
            cudaMemset(p, 0x3c, (size_t)W * H * t.bpp);
        allocations.push_back(p);
        nrd::ResourceCuda res;
        res.data = p; res.width = W; res.height = H; res.pitchBytes = W * t.bpp; res.format = t.format;
        snapshot.SetResource(t.slot, res);
    }
    nrd::CommonSettings common = {};
    const float proj[16] = {1.0f, 0, 0, 0, 0, 1.7f, 0, 0, 0, 0, 0, 1.0f, 0, 0, 0.1f, 0};
    memcpy(common.viewToClipMatrix, proj, 64);
    memcpy(common.viewToClipMatrixPrev, proj, 64);
    common.resourceSize[0] = common.resourceSizePrev[0] = common.rectSize[0] = common.rectSizePrev[0] = W;
    common.resourceSize[1] = common.resourceSizePrev[1] = common.rectSize[1] = common.rectSizePrev[1] = H;
    common.motionVectorScale[0] = 1.0f / W;
    common.motionVectorScale[1] = 1.0f / H;
    if (denoiser == nrd::Denoiser::SIGMA_SHADOW) {
        nrd::SigmaSettings sigmaSettings = {};
        sigmaSettings.lightDirection[2] = 1.0f;
        NRD.SetDenoiserSettings(id, &sigmaSettings);
    }
    cudaStream_t stream;
    cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    uint32_t frame = 0;
    auto step = [&]() {
        NRD.NewFrame();
        common.frameIndex = frame++;
        NRD.SetCommonSettings(common);
        return NRD.Denoise(&id, 1, stream, snapshot) == nrd::Result::SUCCESS;
    };
    for (int i = 0; i < 20; i++)
        if (!step()) { printf("Denoise: %s\n", NRD.GetLastError()); return -1.0; }
    cudaStreamSynchronize(stream);
    cudaEventRecord(e0, stream);
    const auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < frames; i++) step();
    const auto t1 = std::chrono::steady_clock::now();
    cudaEventRecord(e1, stream);
    cudaStreamSynchronize(stream);
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, e0, e1);
    *hostUs = std::chrono::duration<double, std::micro>(t1 - t0).count() / frames;
    uint32_t cached = 0;
    nrdcuGetGraphStats(NRD.GetContext(), &stats[0], &stats[1], &cached);
    stats[2] = cached;
    NRD.Destroy();
    for (void* p : allocations) cudaFree(p);
    cudaStreamDestroy(stream);
    return (double)ms * 1e3 / frames;
}

int main() {
    using RT = nrd::ResourceType;
    using F = nrd::Format;
    const std::vector<Tex> sigma = {{RT::IN_VIEWZ, F::R32_SFLOAT, 4}, {RT::IN_NORMAL_ROUGHNESS, F::R10_G10_B10_A2_UNORM, 4}, {RT::IN_MV, F::RGBA16_SFLOAT, 8}, {RT::IN_PENUMBRA, F::R16_SFLOAT, 2},
                                    {RT::OUT_SHADOW_TRANSLUCENCY, F::R8_UNORM, 1}};
    const std::vector<Tex> reblur = {{RT::IN_MV, F::RGBA16_SFLOAT, 8}, {RT::IN_NORMAL_ROUGHNESS, F::R10_G10_B10_A2_UNORM, 4}, {RT::IN_VIEWZ, F::R32_SFLOAT, 4},
                                     {RT::IN_DIFF_RADIANCE_HITDIST, F::RGBA16_SFLOAT, 8}, {RT::IN_SPEC_RADIANCE_HITDIST, F::RGBA16_SFLOAT, 8},
                                     {RT::OUT_DIFF_RADIANCE_HITDIST, F::RGBA16_SFLOAT, 8}, {RT::OUT_SPEC_RADIANCE_HITDIST, F::RGBA16_SFLOAT, 8}};
    const int frames = 2000;
    printf("{");
    const struct { const char* name; nrd::Denoiser d; uint16_t w, h; const std::vector<Tex>* t; } cases[] = {
        {"sigma_shadow_512x512", nrd::Denoiser::SIGMA_SHADOW, 512, 512, &sigma}, {"reblur_diffuse_specular_256x144", nrd::Denoiser::REBLUR_DIFFUSE_SPECULAR, 256, 144, &reblur}};
    bool first = true;
    for (const auto& c : cases)
        for (int graph = 0; graph < 2; graph++) {
            double hostUs = 0.0;
            uint64_t stats[3] = {};
            const double us = run(c.d, c.w, c.h, *c.t, NRDCU_DEFAULT_FLAGS | (graph ? NRDCU_FLAG_CUDA_GRAPH : 0u), frames, &hostUs, stats);
            if (us < 0.0) return 1;
            printf("%s\"%s%s\": {\"us_per_frame\": %.2f, \"host_us_per_denoise_call\": %.2f, \"graph_captures\": %llu, \"graph_replays\": %llu, \"graphs_cached\": %llu}", first ? "" : ", ", c.name,
                   graph ? "_cuda_graph" : "", us, hostUs, (unsigned long long)stats[0], (unsigned long long)stats[1], (unsigned long long)stats[2]);
            first = false;
        }
    printf(", \"frames\": %d}\n", frames);
    return 0;
}
