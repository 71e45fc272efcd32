// TEST INFRASTRUCTURE — CPU oracle (see hlsl_like.h). SIGMA_SHADOW passes — placeholder until restated.
#include <string>

#include "nrd_shared.h"

namespace orc {
int sigmaDispatch(const std::string&, const void*, uint32_t, Tex*, uint32_t, int, int) { return 1; }
}  // namespace orc
