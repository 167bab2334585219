// tcgen05 TF32 GEMM (placeholder until the tensor-core kernel lands): reports "not eligible".
#include "common.cuh"
namespace avsr {
int gemm_tc(cudaStream_t, int, int, int, int, int, const float*, int, const float*, int, float*, int, float,
            const float*) {
  return -1;
}
}  // namespace avsr
