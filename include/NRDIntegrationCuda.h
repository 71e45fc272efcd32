// NRDIntegrationCuda.h — the CUDA twin of nrd::Integration ( External/NRD/Integration/NRDIntegration.h:211-277 ), header only, over the C ABI of
// nrdcu.h. Same method names, argument meaning and call order as the reference's NRI-based integration, so an application that drives
//     Recreate -> ( per frame: NewFrame, SetCommonSettings, SetDenoiserSettings, Denoise ) -> Destroy
// swaps `nrd::Integration` for `nrd::IntegrationCuda` and hands over CUDA device pointers ( its own, or D3D12 / Vulkan textures imported through
// cudaImportExternalMemory, see INTEGRATION.md ) instead of nri::Texture handles. What differs, because CUDA differs:
//   * `Recreate` takes a CUDA device ordinal where the reference takes an nri::Device;
//   * `Denoise` takes a cudaStream_t ( as void* ) where the reference takes an nri::CommandBuffer: work is enqueued on that stream, in order;
//   * a `Resource` is a pitch-linear device texture ( pointer, size, pitch, nrd::Format ); there are no resource states to track, so
//     `ResourceSnapshot::restoreInitialState` and `Resource::state` do not exist, `userArg` is kept for the application's own bookkeeping;
//   * descriptors and pipelines do not exist: `DestroyCachedDescriptors` and `RecreatePipelines` are kept as no-ops that report success.
// Errors: the reference asserts; this twin returns nrd::Result and keeps the message in `GetLastError()`.
#pragma once

#include <array>
#include <cstdint>
#include <cstring>

#include "nrd_b200.h"
#include "nrdcu.h"

namespace nrd {

struct ResourceCuda {
    void* data = nullptr;      // device pointer to texel ( 0, 0 )
    uint32_t width = 0, height = 0;
    uint32_t pitchBytes = 0;   // multiple of the texel size
    Format format = Format::MAX_NUM;
    void* userArg = nullptr;   // unused by the integration ( NRDIntegration.h:111-113 )
};

// One entry per ResourceType slot ( NRDIntegration.h:121-167 ). The same texture may serve several slots.
struct ResourceSnapshotCuda {
    std::array<ResourceCuda, (size_t)ResourceType::MAX_NUM - 2> slots = {};
    inline void SetResource(ResourceType slot, const ResourceCuda& resource) { slots[(size_t)slot] = resource; }
};

struct IntegrationCudaCreationDesc {
    char name[64] = "";
    uint16_t resourceWidth = 0;    // NRDIntegration.h:184-185
    uint16_t resourceHeight = 0;
    uint32_t flags = NRDCU_DEFAULT_FLAGS;
};

// Threadsafe: no ( like the reference, NRDIntegration.h:210 )
struct IntegrationCuda {
    inline IntegrationCuda() {}
    inline ~IntegrationCuda() { Destroy(); }

    // Creation and re-creation, aka resize. "Destroy" is called under the hood ( NRDIntegration.h:221 )
    inline Result Recreate(const IntegrationCudaCreationDesc& integrationDesc, const InstanceCreationDesc& instanceCreationDesc, int cudaDevice) {
        Destroy();
        m_Desc = integrationDesc;
        const uint32_t r = nrdcuCreate(&instanceCreationDesc, integrationDesc.resourceWidth, integrationDesc.resourceHeight, cudaDevice, integrationDesc.flags, &m_Context);
        if (r != 0) m_Context = nullptr;
        return (Result)r;
    }

    // Must be called once on a frame start ( NRDIntegration.h:233 ). The reference rotates its constant-buffer ring here; constants travel as kernel
    // parameters in this executor, so only the frame counter advances.
    inline void NewFrame() { m_FrameIndex++; }

    inline Result SetCommonSettings(const CommonSettings& commonSettings) { return m_Context ? (Result)nrdcuSetCommonSettings(m_Context, &commonSettings) : Result::FAILURE; }
    inline Result SetDenoiserSettings(Identifier denoiser, const void* denoiserSettings) {
        return m_Context ? (Result)nrdcuSetDenoiserSettings(m_Context, denoiser, denoiserSettings) : Result::FAILURE;
    }

    // Invoke denoising for the specified denoisers on `cudaStream` ( NRDIntegration.h:243 ). Every slot the dispatches reference must be set.
    inline Result Denoise(const Identifier* denoisers, uint32_t denoisersNum, void* cudaStream, const ResourceSnapshotCuda& resourceSnapshot) {
        if (!m_Context) return Result::FAILURE;
        for (size_t slot = 0; slot < resourceSnapshot.slots.size(); slot++) {
            const ResourceCuda& r = resourceSnapshot.slots[slot];
            if (!r.data) continue;
            const nrdcuTexture t = {r.data, r.width, r.height, r.pitchBytes, (uint32_t)r.format};
            const uint32_t rc = nrdcuSetResource(m_Context, (uint32_t)slot, &t);
            if (rc != 0) return (Result)rc;
        }
        return (Result)nrdcuDenoise(m_Context, denoisers, denoisersNum, cudaStream);
    }

    inline void Destroy() {
        if (m_Context) nrdcuDestroy(m_Context);   // waits for the device to go idle ( "autoWaitForIdle", NRDIntegration.h:199-200 )
        m_Context = nullptr;
        m_FrameIndex = 0;
    }

    // Kept for source compatibility: there are no descriptors or pipelines to rebuild ( kernels are resolved when the library loads )
    inline void DestroyCachedDescriptors() {}
    inline bool RecreatePipelines() { return true; }

    // (Optional) Statistics ( NRDIntegration.h:266-277 )
    inline double GetTotalMemoryUsageInMb() const { return GetPersistentMemoryUsageInMb() + GetAliasableMemoryUsageInMb(); }
    inline double GetPersistentMemoryUsageInMb() const { return usage(0); }
    inline double GetAliasableMemoryUsageInMb() const { return usage(1); }
    inline double GetPrivateMemoryUsageInMb() const { return usage(2); }   // not in the reference: the executor's own scratch ( geometry plane )

    inline const char* GetLastError() const { return nrdcuGetLastError(); }
    inline nrdcuContext* GetContext() const { return m_Context; }   // for the nrdcu* calls without a counterpart ( strips, host frames, profiling )
    inline uint32_t GetFrameIndex() const { return m_FrameIndex; }

private:
    IntegrationCuda(const IntegrationCuda&) = delete;
    inline double usage(int which) const {
        uint64_t v[3] = {0, 0, 0};
        if (m_Context) nrdcuGetMemoryUsage(m_Context, &v[0], &v[1], &v[2]);
        return double(v[which]) / (1024.0 * 1024.0);
    }

    IntegrationCudaCreationDesc m_Desc = {};
    nrdcuContext* m_Context = nullptr;
    uint32_t m_FrameIndex = 0;
};

}  // namespace nrd
