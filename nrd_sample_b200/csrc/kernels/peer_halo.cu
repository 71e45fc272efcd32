// Seam traffic of the multi-GPU strip mode over NVLink peer memory (include/nrdcu.h "strips over peer memory").
//
// After a pass has written the rows of its strip, the rows next to the seams are stored straight into the neighbour
// GPU's copy of the same texture (its apron) through a peer mapping of that allocation (CUDA IPC), followed by a
// release-store of the pass sequence number into a flag word in the neighbour's memory; before the next pass starts
// each GPU spins (one thread) on its own flag words until both neighbours have delivered. No host round trip, no
// collective library: three tiny launches per pass on the stream that runs the denoiser.
#include <cuda_runtime.h>

#include <cstdint>

#include "peer_halo.cuh"

namespace nrdk {

namespace {

// One grid row per segment; 16-byte vector loads from local HBM, 16-byte stores over NVLink
__global__ void __launch_bounds__(256) haloPushKernel(const __grid_constant__ HaloSegments segs) {
    const HaloSegment g = segs.s[blockIdx.y];
    const uint4* __restrict__ src = reinterpret_cast<const uint4*>(g.src);
    uint4* __restrict__ dst = reinterpret_cast<uint4*>(g.dst);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < g.vecs; i += gridDim.x * blockDim.x) dst[i] = src[i];
}

__global__ void haloSignalKernel(uint32_t* slotInUpNeighbour, uint32_t* slotInDownNeighbour, uint32_t seq) {
    // stream order already put every store of the pass and of the push kernel before this launch; the fence + release make
    // them visible system-wide before the flag is
    __threadfence_system();
    if (slotInUpNeighbour) asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(slotInUpNeighbour), "r"(seq) : "memory");
    if (slotInDownNeighbour) asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(slotInDownNeighbour), "r"(seq) : "memory");
}

__global__ void haloWaitKernel(const uint32_t* slots, uint32_t seq, int waitUp, int waitDown, volatile uint32_t* hostError, long long timeoutCycles) {
    const long long t0 = clock64();
    for (int k = 0; k < 2; k++) {
        if (!(k == 0 ? waitUp : waitDown)) continue;
        for (;;) {
            uint32_t v;
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(slots + k) : "memory");
            if ((int32_t)(v - seq) >= 0) break;
            if (clock64() - t0 > timeoutCycles) {  // a neighbour died or fell out of lockstep: report instead of hanging the GPU
                *hostError = 1u + (uint32_t)k;
                return;
            }
            __nanosleep(200);
        }
    }
}

}  // namespace

void launchHaloPush(const HaloSegments& segs, cudaStream_t stream) {
    if (!segs.n) return;
    haloPushKernel<<<dim3(32, segs.n), 256, 0, stream>>>(segs);
}
void launchHaloSignal(uint32_t* slotInUpNeighbour, uint32_t* slotInDownNeighbour, uint32_t seq, cudaStream_t stream) {
    haloSignalKernel<<<1, 1, 0, stream>>>(slotInUpNeighbour, slotInDownNeighbour, seq);
}
void launchHaloWait(const uint32_t* slots, uint32_t seq, bool waitUp, bool waitDown, uint32_t* hostError, cudaStream_t stream) {
    // ~4 s at 1.9 GHz: far beyond any frame, short enough not to look like a hung GPU to the caller
    haloWaitKernel<<<1, 1, 0, stream>>>(slots, seq, waitUp ? 1 : 0, waitDown ? 1 : 0, hostError, 8000000000ll);
}

}  // namespace nrdk
