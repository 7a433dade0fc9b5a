// tcgen05/TMEM tensor-core path -- placeholder until the kernel lands.
#pragma once
#include <string>
#include <vector>
#include "common.cuh"

namespace tck {
struct Plan { int dummy; };
inline bool build_plan(int, const int*, const float* const*, const float* const*, const int*, Plan&,
                       std::vector<unsigned short>*, std::vector<float>&, std::string& why) {
    why = "tensor-core kernel not built yet";
    return false;
}
inline cudaError_t prepare() { return cudaSuccess; }
inline cudaError_t launch(const Plan&, const NormConsts&, const LaunchArgs&, const void*, const float*, int, int,
                          cudaStream_t) {
    return cudaErrorNotSupported;
}
}  // namespace tck
