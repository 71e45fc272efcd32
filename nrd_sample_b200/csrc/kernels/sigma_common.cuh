// SIGMA helpers and launch-parameter blocks (filled in with the SIGMA kernels).
#pragma once
#include "common.cuh"
