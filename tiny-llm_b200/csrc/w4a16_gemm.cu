// W4A16 prefill GEMM (tcgen05) - placeholder: the tensor-core kernel is not
// built yet, so w4a16_gemm_supported() reports false and tl_quantized_matmul
// routes every M through the weight-streaming kernel in 8..32-row passes.
#include "common.cuh"
#include "kernels.h"

namespace tl {

bool w4a16_gemm_supported(int, int, int, int) { return false; }
int w4a16_gemm_split(int, int, int, int) { return 1; }
size_t w4a16_gemm_workspace(int, int, int, int, int) { return 0; }
int launch_w4a16_gemm(const void *, const void *, const void *, const void *, void *, int, int, int, int, int, void *,
                      size_t, cudaStream_t) {
    return fail(TL_EINVAL, "quantized_matmul: tcgen05 GEMM not available in this build");
}

}  // namespace tl
