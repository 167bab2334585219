// Micro-test: tcgen05.mma with the A operand in tensor memory (".ts" form).  Checks the assumed TMEM layout of a
// K-major fp16 A tile (lane = row m, 32-bit column c holds {A[m][2c], A[m][2c+1]}) against a host product.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/micro/ts_mma_test tools/micro/ts_mma_test.cu
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>

constexpr int M = 128, N = 32, K = 64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc_k128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

__global__ void __launch_bounds__(160, 1) ts_kernel(const __half* A, const __half* Bm, float* D) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sB = base, sBar = base + N * 128, sTmem = sBar + 8;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sBar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 4) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(sTmem) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // B operand: N rows x 64 halves, SWIZZLE_128B
  for (int i = tid; i < N * K; i += blockDim.x) {
    const int n = i / K, k = i % K;
    const uint32_t off = n * 128 + ((((k >> 3) ^ (n & 7)) << 4)) + ((k & 7) << 1);
    *reinterpret_cast<__half*>(gen + off) = Bm[i];
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(sTmem));
  const uint32_t tA = tmem_base + 64;  // A tile at columns 64..95, D at columns 0..31
  if (warp < 4) {
    const int m = 32 * warp + lane;
    uint32_t r[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) {
      __half2 h = __halves2half2(A[m * K + 2 * c], A[m * K + 2 * c + 1]);
      r[c] = *reinterpret_cast<uint32_t*>(&h);
    }
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(tA + ((uint32_t)(32 * warp) << 16)),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
        "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
        "r"(r[31])
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (warp == 4 && lane == 0) {
#pragma unroll
    for (int k4 = 0; k4 < K / 16; ++k4) {
      const uint64_t db = make_desc_k128(sB + k4 * 32);
      const uint32_t acc = k4 ? 1u : 0u;
      asm volatile(
          "{\n\t"
          ".reg .pred p;\n\t"
          "setp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
          "}" ::"r"(tmem_base),
          "r"(tA + k4 * 8), "l"(db), "r"(IDESC), "r"(acc)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(sBar) : "memory");
  }
  if (warp < 4) {
    asm volatile(
        "{\n\t.reg .pred P1;\nW:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0;\n\t@P1 bra DONE;\n\tbra W;\nDONE:\n\t}" ::"r"(sBar)
        : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(tmem_base + ((uint32_t)(32 * warp) << 16)));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    const int m = 32 * warp + lane;
    for (int n = 0; n < N; ++n) D[m * N + n] = __uint_as_float(r[n]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 4) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 128;" ::"r"(tmem_base) : "memory");
}

int main() {
  std::vector<__half> A(M * K), B(N * K);
  std::vector<float> Af(M * K), Bf(N * K), ref(M * N), out(M * N);
  srand(1);
  for (int i = 0; i < M * K; ++i) { A[i] = __float2half((rand() % 2001 - 1000) / 1000.0f); Af[i] = __half2float(A[i]); }
  for (int i = 0; i < N * K; ++i) { B[i] = __float2half((rand() % 2001 - 1000) / 1000.0f); Bf[i] = __half2float(B[i]); }
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double s = 0;
      for (int k = 0; k < K; ++k) s += (double)Af[m * K + k] * Bf[n * K + k];
      ref[m * N + n] = (float)s;
    }
  __half *dA, *dB; float* dD;
  cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dB, B.size() * 2); cudaMalloc(&dD, out.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0, out.size() * 4);
  ts_kernel<<<1, 160, N * 128 + 64 + 1024>>>(dA, dB, dD);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  cudaMemcpy(out.data(), dD, out.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0;
  for (int i = 0; i < M * N; ++i) maxerr = fmax(maxerr, fabs(out[i] - ref[i]));
  printf("TS-MMA max abs err vs host = %.3e  (ref[0]=%.4f out[0]=%.4f ref[last]=%.4f out[last]=%.4f)\n", maxerr, ref[0], out[0],
         ref[M * N - 1], out[M * N - 1]);
  printf(maxerr < 1e-3 ? "TS_MMA_OK\n" : "TS_MMA_MISMATCH\n");
  return 0;
}
