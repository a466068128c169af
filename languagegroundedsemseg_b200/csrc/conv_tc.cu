// tcgen05 path — placeholder until the tensor-core kernel lands (returns UNSUPPORTED so the SIMT kernel runs).
#include "common.cuh"
namespace lgs {
bool tc_built() { return false; }
int conv_fwd_tc(const void*, int64_t, int, const void*, int, int, const int32_t*, int64_t, int, const float*, void*,
                int, cudaStream_t) { return LGS_E_UNSUPPORTED; }
int conv_wgrad_tc(const void*, int64_t, int, const void*, int64_t, int, const int32_t*, int, float*, int,
                  cudaStream_t) { return LGS_E_UNSUPPORTED; }
}  // namespace lgs
