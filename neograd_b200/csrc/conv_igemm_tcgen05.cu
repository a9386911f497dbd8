// Implicit-GEMM convolution on tcgen05 (placeholder dispatcher: filled in by the TMA-fed kernel).
#include "tc_common.cuh"
namespace ngb {
int conv_igemm_eligible(int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t) { return 0; }
int conv_igemm_fprop(const float*, const float*, const float*, float*, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t,
                     int64_t, int64_t, cudaStream_t) { return fail(NGB_ERR_UNSUPPORTED, "conv igemm not built"); }
int conv_igemm_dgrad(const float*, const float*, float*, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t,
                     int64_t, int, cudaStream_t) { return fail(NGB_ERR_UNSUPPORTED, "conv igemm not built"); }
}  // namespace ngb
