// SIGMA_SHADOW passes on sm_100a — not implemented yet; the executor reports UNSUPPORTED for these shaders.
#include <string>

#include "../../../include/nrd_b200.h"
#include "../../../include/nrdcu.h"
#include "sigma_common.cuh"

namespace nrdk {
uint32_t dispatchSigma(const std::string& id, const void*, uint32_t, const nrdcuTexture*, uint32_t, cudaStream_t, std::string& err) {
    err = "no CUDA kernel for shader '" + id + "' yet";
    return (uint32_t)nrd::Result::UNSUPPORTED;
}
}  // namespace nrdk
