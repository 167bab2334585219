// L2 read bandwidth of the GPU: the denominator for the attention sweeps, which read L2-resident fp16 memories.
// Every CTA streams its slice of an S MB buffer REPS times inside one launch (16-byte loads, 8 in flight per thread).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/l2_bw tools/l2_bw.cu && tools/l2_bw
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256) read_kernel(const uint4* __restrict__ p, long long n16, int reps, unsigned* sink) {
  unsigned acc = 0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (int r = 0; r < reps; ++r) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 7 * stride < n16; i += 8 * stride) {
      uint4 v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = __ldcg(p + i + k * stride);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc += v[k].x ^ v[k].y ^ v[k].z ^ v[k].w;
    }
    for (; i < n16; i += stride) {
      uint4 v = __ldcg(p + i);
      acc += v.x ^ v.y ^ v.z ^ v.w;
    }
  }
  if (acc == 0x12345678u) *sink = acc;
}
int main() {
  unsigned* sink;
  cudaMalloc(&sink, 4);
  for (int mb : {8, 20, 40, 80, 110, 512}) {
    const long long bytes = (long long)mb << 20, n16 = bytes / 16;
    uint4* buf;
    cudaMalloc(&buf, bytes);
    cudaMemset(buf, 1, bytes);
    for (int ctas_per_sm : {1, 2, 4}) {
      const int grid = 148 * ctas_per_sm, reps = mb <= 128 ? 200 : 8;
      read_kernel<<<grid, 256>>>(buf, n16, 2, sink);
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0);
      read_kernel<<<grid, 256>>>(buf, n16, reps, sink);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      printf("%4d MB, %d CTAs/SM x 256 threads: %8.1f GB/s\n", mb, ctas_per_sm, (double)bytes * reps / (ms * 1e-3) / 1e9);
    }
    cudaFree(buf);
  }
  return 0;
}
