// Seam traffic of the multi-GPU strip mode over NVLink peer memory (include/nrdcu.h "strips over peer memory").
//
// After a pass has written the rows of its strip, the rows next to the seams are stored straight into the neighbour
// GPU's copy of the same texture (its apron) through a peer mapping of that allocation (CUDA IPC), followed by a
// release-store of the pass sequence number into a flag word in the neighbour's memory; before the next pass starts
// each GPU spins (one thread) on its own flag words until both neighbours have delivered. No host round trip, no
// collective library: three tiny launches per pass on the stream that runs the denoiser.
#include <cuda_fp16.h>
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

// KIND: 0 = RGBA16F, 1 = RG16F, 2 = RGBA32F, 3 = RG32F — only .y is read
template <int KIND>
__global__ void __launch_bounds__(256) motionBoundKernel(const uint8_t* __restrict__ mv, uint32_t pitchBytes, int width, int y0, int y1, int frameHeight, float scaleYRows, int bound,
                                                         uint32_t* worstExcessRows) {
    const int px = blockIdx.x * 32 + (threadIdx.x & 31), py = y0 + blockIdx.y * 8 + (threadIdx.x >> 5);
    float excess = 0.0f;
    if (px < width && py < y1) {
        const uint8_t* row = mv + (size_t)py * pitchBytes;
        float my;
        if (KIND == 0) my = __half2float(__ushort_as_half(__ldg(reinterpret_cast<const unsigned short*>(row) + px * 4 + 1)));
        else if (KIND == 1) my = __half2float(__ushort_as_half(__ldg(reinterpret_cast<const unsigned short*>(row) + px * 2 + 1)));
        else if (KIND == 2) my = __ldg(reinterpret_cast<const float*>(row) + px * 4 + 1);
        else my = __ldg(reinterpret_cast<const float*>(row) + px * 2 + 1);
        const float prevRow = (float)py + 0.5f + my * scaleYRows;
        if (prevRow >= 0.0f && prevRow < (float)frameHeight) {   // history outside the screen is not fetched
            const float lo = (float)(y0 - bound), hi = (float)(y1 + bound);
            excess = fmaxf(lo - prevRow, prevRow - hi);
        }
    }
    const uint32_t rows = excess > 0.0f ? (uint32_t)ceilf(excess) : 0u;
    const uint32_t warpWorst = __reduce_max_sync(0xFFFFFFFFu, rows);
    if (warpWorst && (threadIdx.x & 31) == 0) atomicMax_system(worstExcessRows, warpWorst);
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

namespace nrdk {
bool launchMotionBoundCheck(const void* mv, uint32_t pitchBytes, uint32_t format, int width, int y0, int y1, int frameHeight, float scaleYRows, int bound, uint32_t* worstExcessRows,
                            cudaStream_t stream) {
    if (y1 <= y0 || !mv) return false;
    const dim3 grid((width + 31) / 32, (y1 - y0 + 7) / 8);
    const uint8_t* p = (const uint8_t*)mv;
    switch (format) {
        case 27: motionBoundKernel<0><<<grid, 256, 0, stream>>>(p, pitchBytes, width, y0, y1, frameHeight, scaleYRows, bound, worstExcessRows); return true;   // RGBA16_SFLOAT
        case 22: motionBoundKernel<1><<<grid, 256, 0, stream>>>(p, pitchBytes, width, y0, y1, frameHeight, scaleYRows, bound, worstExcessRows); return true;   // RG16_SFLOAT
        case 39: motionBoundKernel<2><<<grid, 256, 0, stream>>>(p, pitchBytes, width, y0, y1, frameHeight, scaleYRows, bound, worstExcessRows); return true;   // RGBA32_SFLOAT
        case 33: motionBoundKernel<3><<<grid, 256, 0, stream>>>(p, pitchBytes, width, y0, y1, frameHeight, scaleYRows, bound, worstExcessRows); return true;   // RG32_SFLOAT
        default: return false;
    }
}
}  // namespace nrdk
