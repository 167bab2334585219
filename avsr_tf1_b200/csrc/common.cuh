// Shared helpers for the avsr_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

namespace avsr {

// ---- error reporting across the C ABI (no exceptions) ----------------------
void set_error(const char* fmt, ...);
extern unsigned long long g_launch_count;  // kernels launched by this library

#define AVSR_CHECK_CUDA(expr)                                                          \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess) {                                                           \
      ::avsr::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return 1;                                                                        \
    }                                                                                  \
  } while (0)

#define AVSR_REQUIRE(cond, ...)                                                        \
  do {                                                                                 \
    if (!(cond)) {                                                                     \
      ::avsr::set_error(__VA_ARGS__);                                                  \
      return 2;                                                                        \
    }                                                                                  \
  } while (0)

// launch + count + check (no sync)
#define AVSR_LAUNCH(kernel, grid, block, smem, stream, ...)                            \
  do {                                                                                 \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                        \
    ++::avsr::g_launch_count;                                                          \
    AVSR_CHECK_CUDA(cudaGetLastError());                                               \
  } while (0)

#define AVSR_TRY(expr)                                                                 \
  do {                                                                                 \
    int _r = (expr);                                                                   \
    if (_r != 0) return _r;                                                            \
  } while (0)

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- optional device timing of the persistent kernels (avsr_kernel_timing) ----
// Kernel classes; when timing is enabled a launcher brackets its kernel with CUDA events on the launching stream
// (never under stream capture) and the elapsed times are summed per class.
enum { AVSR_K_ATTN_FWD = 0, AVSR_K_ATTN_BWD = 1, AVSR_K_LSTM_FWD = 2, AVSR_K_LSTM_BWD = 3, AVSR_K_GEMM = 4, AVSR_K_COUNT = 5 };
bool kernel_timing_enabled();
// returns an opaque slot (>= 0) after recording the start event, or -1 when timing is off
int kernel_timer_begin(cudaStream_t st, int klass);
void kernel_timer_end(cudaStream_t st, int slot);

// ---- device helpers ---------------------------------------------------------
__device__ __forceinline__ float sigmoidf_acc(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
// tanh via exp: accurate to ~1e-7 relative (tanh.approx is only ~5e-4)
__device__ __forceinline__ float tanhf_acc(float x) {
  float ax = fabsf(x);
  float e = __expf(-2.0f * ax);
  float r = __fdividef(1.0f - e, 1.0f + e);
  return copysignf(r, x);
}

// round-to-nearest fp32 -> tf32 (low 13 mantissa bits zero).  Operands of the tcgen05 kind::tf32 products
// are rounded where they are PRODUCED, so the tensor core's own truncation is a no-op and the products
// carry unbiased rounding error instead of a systematic -2^-11 shrink.
__device__ __forceinline__ float tf32_rn(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ float maybe_tf32(float x, int rnd) { return rnd ? tf32_rn(x) : x; }

// ---- counter-based random words (dropout masks, scheduled sampling) -----------------------------------------------
// Two rounds of the murmur3 32-bit finaliser over (seed, step, stream, hi, lo); no state, so forward and backward
// regenerate the same mask.  Restated bit for bit in oracle/avsr_oracle.py (rand_u32).
__host__ __device__ __forceinline__ uint32_t fmix32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x85EBCA6Bu;
  x ^= x >> 13;
  x *= 0xC2B2AE35u;
  x ^= x >> 16;
  return x;
}
__host__ __device__ __forceinline__ uint32_t avsr_rand_u32(uint32_t seed, uint32_t step, uint32_t stream, uint32_t hi,
                                                          uint32_t lo) {
  uint32_t x = fmix32(lo * 0x9E3779B1u + seed);
  x += hi * 0x27D4EB2Fu + stream * 0x165667B1u + step * 0x9E3779B9u;
  return fmix32(x ^ 0x5BD1E995u);
}
// inverted-dropout factor of element (hi, lo): 1/keep if kept, else 0.  thr == 0: dropout off (factor 1).
__device__ __forceinline__ float drop_factor(const uint32_t* __restrict__ rng, uint32_t stream, uint32_t thr,
                                             float inv_keep, uint32_t hi, uint32_t lo) {
  if (thr == 0u) return 1.0f;
  return avsr_rand_u32(rng[0], rng[1], stream, hi, lo) < thr ? inv_keep : 0.0f;
}
static inline float inv_keep_of(uint32_t thr) { return thr ? (float)(4294967296.0 / (double)thr) : 1.0f; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// block-wide reductions for blockDim.x <= 1024 (multiple of 32); `sh` has >= 33 floats
__device__ __forceinline__ float block_sum(float v, float* sh) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  if (w == 0) {
    float t = lane < nw ? sh[lane] : 0.0f;
    t = warp_sum(t);
    if (lane == 0) sh[32] = t;
  }
  __syncthreads();
  return sh[32];
}
__device__ __forceinline__ float block_max(float v, float* sh) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) sh[w] = v;
  __syncthreads();
  if (w == 0) {
    float t = lane < nw ? sh[lane] : -INFINITY;
    t = warp_max(t);
    if (lane == 0) sh[32] = t;
  }
  __syncthreads();
  return sh[32];
}

// ---- internal entry points shared between translation units ------------------
// C[M,N](ldc) = beta*C + op(A)*op(B) (+bias[N]); beta in {0,1}.  op(A) is MxK:
// transA==0 -> A stored [M,K] row-major with leading dim lda; transA==1 -> stored [K,M].
// round_out: store C rounded to tf32 (only with beta == 0; used when C is itself a tensor-core operand).
int gemm(cudaStream_t st, int transA, int transB, int M, int N, int K, const float* A, int lda, const float* B,
         int ldb, float* C, int ldc, float beta, const float* bias, int round_out = 0);
// exact fp32 CUDA-core implementation (gemm_simt.cu)
int gemm_simt(cudaStream_t st, int transA, int transB, int M, int N, int K, const float* A, int lda, const float* B,
              int ldb, float* C, int ldc, float beta, const float* bias, int round_out = 0);
// 1 when the tcgen05 TF32 path (and producer-side tf32 rounding) is enabled
int tensor_cores_enabled();

}  // namespace avsr
