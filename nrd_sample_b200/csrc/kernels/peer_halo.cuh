// Launch interface of kernels/peer_halo.cu (seam rows pushed into the neighbour GPU's textures over NVLink peer memory).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace nrdk {

struct HaloSegment {
    const void* src;  // first byte of the rows in this GPU's texture
    void* dst;        // the same rows in the neighbour's copy (peer mapping)
    uint32_t vecs;    // length in 16-byte units (rows x pitch; pitches are multiples of 256)
    uint32_t _pad;
};
constexpr uint32_t kMaxHaloSegments = 24;  // <= 11 written textures per pass (RELAX temporal accumulation) x 2 directions
struct HaloSegments {
    HaloSegment s[kMaxHaloSegments];
    uint32_t n;
};

// Strips read the previous frame at pixel + motion: rows of vertical motion beyond the apron would be read from a neighbour's rows this GPU never received.
// One pass over the strip's motion vectors ( 2D / 2.5D, uv space after `scaleYRows` = motionVectorScale.y x rect height ) raises *worstExcessRows ( pinned,
// mapped ) to the largest number of rows any in-frame history fetch lands beyond [ y0 - bound, y1 + bound ). `format`: nrd::Format of the texture
// ( RGBA16_SFLOAT, RG16_SFLOAT, RGBA32_SFLOAT, RG32_SFLOAT ); returns false for anything else ( nothing launched ).
bool launchMotionBoundCheck(const void* mv, uint32_t pitchBytes, uint32_t format, int width, int y0, int y1, int frameHeight, float scaleYRows, int bound, uint32_t* worstExcessRows,
                            cudaStream_t stream);
void launchHaloPush(const HaloSegments& segs, cudaStream_t stream);
void launchHaloSignal(uint32_t* slotInUpNeighbour, uint32_t* slotInDownNeighbour, uint32_t seq, cudaStream_t stream);
void launchHaloWait(const uint32_t* slots, uint32_t seq, bool waitUp, bool waitDown, uint32_t* hostError, cudaStream_t stream);

}  // namespace nrdk
