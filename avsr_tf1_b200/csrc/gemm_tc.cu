// tcgen05 TF32 GEMM for sm_100a: TMA (cp.async.bulk.tensor, 128B swizzle) -> shared memory
// -> tcgen05.mma.kind::tf32 (fp32 operands read as TF32, fp32 accumulators in TMEM) ->
// tcgen05.ld -> swizzled staging tile -> TMA store / TMA reduce-add.
//
// C[M,N](ldc) = beta*C + op(A) op(B) (+bias),  beta in {0,1}.
//
// Operand "major-ness" is taken straight from the caller's layout, so no transposes are
// materialised:  A [M,K] row-major -> K-major;  A stored [K,M] -> MN-major;
//                B stored [N,K]    -> K-major;  B [K,N] row-major -> MN-major.
// Tiles: BLOCK_M = 128 (one UMMA M), BLOCK_N in {32..256} chosen per problem, BLOCK_K = 32
// fp32 (one 128-byte swizzle row).  Persistent CTAs loop over (m-tile, n-tile, k-split) work
// items; warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-5 = epilogue; the accumulator
// is double-buffered in TMEM (2 x 256 columns) so the epilogue of one tile overlaps the
// main loop of the next.  beta = 1 and split-K use the TMA reduce-add (no read of C).
#include <cuda.h>

#include "common.cuh"

namespace avsr {

namespace tc {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 32;            // fp32 elements = 128 bytes = one swizzle row
constexpr int UMMA_K = 8;              // tf32
constexpr int STAGES = 4;
constexpr int A_BYTES = BLOCK_M * BLOCK_K * 4;  // 16 KB
constexpr int B_BYTES_MAX = 256 * BLOCK_K * 4;  // 32 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES_MAX;
constexpr int EPI_TILE_BYTES = 32 * 32 * 4;     // 4 KB per warp per buffer
constexpr int EPI_BYTES = 4 * 2 * EPI_TILE_BYTES;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 256 + 1024;  // + barriers + alignment slack
constexpr int THREADS = 192;

struct Params {
  int M, N, K;
  int block_n, tiles_m, tiles_n, splitk, k_tiles, k_tiles_per_split, num_work;
  int a_mn, b_mn, reduce, round_out;
  const float* bias;
  uint32_t idesc;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t"
      "}" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(tm), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(tm), "r"(src),
               "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* tm, uint32_t src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(tm),
               "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor), version 1.
// layout_type: 2 = SWIZZLE_128B (K-major tiles), 1 = SWIZZLE_128B_BASE32B - the only layout tcgen05
// accepts for MN-major TF32 operands (128-byte rows, 32-byte chunks XOR-ed with row & 3; TMA produces
// it with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                              uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // version = 1 (Blackwell)
  d |= (uint64_t)layout_type << 61;
  return d;
}

__global__ void __launch_bounds__(THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmC, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B tiles need 1024 B alignment
  const uint32_t stage0 = base;
  const uint32_t epi0 = base + STAGES * STAGE_BYTES;
  const uint32_t bars = epi0 + EPI_BYTES;
  // barrier map (8 B each): full[STAGES] empty[STAGES] tmem_full[2] tmem_empty[2]; then tmem base slot
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bars + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bars + 8u * (2 * STAGES + 2 + a); };
  const uint32_t tmem_slot = bars + 8u * (2 * STAGES + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmC) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {  // whole warp allocates all 512 TMEM columns (1 CTA / SM)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(tmem_slot) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  const int b_bytes = p.block_n * BLOCK_K * 4;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = blockIdx.x; w < p.num_work; w += gridDim.x) {
        const int ks = w % p.splitk;
        const int t = w / p.splitk;
        const int tn = t % p.tiles_n, tm = t / p.tiles_n;
        const int m0 = tm * BLOCK_M, n0 = tn * p.block_n;
        const int kt0 = ks * p.k_tiles_per_split;
        const int kt1 = min(p.k_tiles, kt0 + p.k_tiles_per_split);
        for (int kt = kt0; kt < kt1; ++kt) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          mbar_expect_tx(full_bar(stage), A_BYTES + b_bytes);
          const uint32_t sa = stage0 + stage * STAGE_BYTES, sb = sa + A_BYTES;
          const int k0 = kt * BLOCK_K;
          if (!p.a_mn) {
            tma_load_2d(sa, &tmA, full_bar(stage), k0, m0);
          } else {
#pragma unroll
            for (int i = 0; i < BLOCK_M / 32; ++i) tma_load_2d(sa + i * 4096, &tmA, full_bar(stage), m0 + 32 * i, k0);
          }
          if (!p.b_mn) {
            tma_load_2d(sb, &tmB, full_bar(stage), k0, n0);
          } else {
            for (int i = 0; i < p.block_n / 32; ++i) tma_load_2d(sb + i * 4096, &tmB, full_bar(stage), n0 + 32 * i, k0);
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int w = blockIdx.x; w < p.num_work; w += gridDim.x, ++it) {
      const int ks = w % p.splitk;
      const int kt0 = ks * p.k_tiles_per_split;
      const int kt1 = min(p.k_tiles, kt0 + p.k_tiles_per_split);
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(tempty_bar(acc), acc_phase ^ 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t tmem_d = tmem_base + acc * 256;
      for (int kt = kt0; kt < kt1; ++kt) {
        mbar_wait(full_bar(stage), phase);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (lane == 0) {
          const uint32_t sa = stage0 + stage * STAGE_BYTES, sb = sa + A_BYTES;
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // K-major: 8-row groups 1024 B apart (SBO), K advances 32 B inside the swizzle row.
            // MN-major: 32-element MN blocks 4096 B apart (LBO); atoms of 4 K-rows are 512 B apart
            //           (SBO); one UMMA_K = 8 K-rows = 1024 B.
            const uint64_t da = p.a_mn ? make_desc(sa + k * 1024, 4096, 512, 1) : make_desc(sa + k * 32, 16, 1024, 2);
            const uint64_t db = p.b_mn ? make_desc(sb + k * 1024, 4096, 512, 1) : make_desc(sb + k * 32, 16, 1024, 2);
            umma_tf32(tmem_d, da, db, p.idesc, (kt > kt0 || k > 0) ? 1u : 0u);
          }
          umma_commit(empty_bar(stage));            // smem slot free once these MMAs retire
          if (kt == kt1 - 1) umma_commit(tfull_bar(acc));  // accumulator ready for the epilogue
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (kt1 <= kt0 && lane == 0) umma_commit(tfull_bar(acc));  // empty k-range (never for valid launches)
    }
  } else {
    // ================= epilogue warps 2..5 =================
    const int q = warp & 3;  // TMEM lane quadrant this warp may access
    const uint32_t my_epi = epi0 + (warp - 2) * 2 * EPI_TILE_BYTES;
    int it = 0, ebuf = 0;
    for (int w = blockIdx.x; w < p.num_work; w += gridDim.x, ++it) {
      const int ks = w % p.splitk;
      const int t = w / p.splitk;
      const int tn = t % p.tiles_n, tm = t / p.tiles_n;
      const int m0 = tm * BLOCK_M, n0 = tn * p.block_n;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(tfull_bar(acc), acc_phase);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int row0 = m0 + 32 * q;
      const bool rows_live = row0 < p.M;
      for (int c = 0; c < p.block_n / 32; ++c) {
        const int col0 = n0 + 32 * c;
        if (col0 >= p.N) break;
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(32 * q) << 16) + acc * 256 + c * 32, r);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (!rows_live) continue;
        if (p.bias != nullptr && ks == 0) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float bj = (col0 + j < p.N) ? __ldg(p.bias + col0 + j) : 0.0f;
            r[j] = __float_as_uint(__uint_as_float(r[j]) + bj);
          }
        }
        if (p.round_out) {
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(tf32_rn(__uint_as_float(r[j])));
        }
        // the staging buffer we are about to overwrite must have been read by its previous TMA store
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        __syncwarp();
        const uint32_t tile = my_epi + ebuf * EPI_TILE_BYTES;
        const uint32_t rowaddr = tile + lane * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j) {  // 16-byte chunk j of row `lane` lives at chunk (j ^ (lane & 7))
          const uint32_t dst = rowaddr + ((j ^ (lane & 7)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(r[4 * j]), "r"(r[4 * j + 1]),
                       "r"(r[4 * j + 2]), "r"(r[4 * j + 3])
                       : "memory");
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          if (p.reduce) tma_reduce_add_2d(&tmC, tile, col0, row0);
          else tma_store_2d(&tmC, tile, col0, row0);
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        ebuf ^= 1;
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    __syncwarp();
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// ---------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// 2-D fp32 tensor map: inner dimension `d0` (contiguous), outer `d1` with row pitch `ld` elements
static int make_map(CUtensorMap* tm, const float* ptr, uint64_t d0, uint64_t d1, uint64_t ld, uint32_t b0, uint32_t b1,
                    CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return 1;
  cuuint64_t gdim[2] = {d0, d1};
  cuuint64_t gstr[1] = {ld * sizeof(float)};
  cuuint32_t box[2] = {b0, b1};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)ptr, gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : 1;
}

}  // namespace tc

// Returns -1 when the problem is not eligible for the tensor-core path (caller falls back to the
// exact-fp32 CUDA-core kernel: odd leading dimensions such as the N=31 logits, or tiny products).
int gemm_tc(cudaStream_t st, int transA, int transB, int M, int N, int K, const float* A, int lda, const float* B,
            int ldb, float* C, int ldc, float beta, const float* bias, int round_out) {
  using namespace tc;
  if (M <= 0 || N <= 0 || K <= 0) return -1;
  if ((lda & 3) || (ldb & 3) || (ldc & 3)) return -1;
  if (((uintptr_t)A & 15) || ((uintptr_t)B & 15) || ((uintptr_t)C & 15)) return -1;
  if (beta != 0.0f && beta != 1.0f) return -1;
  const double macs = (double)M * N * K;
  if (macs < 32.0 * 1024 * 1024) return -1;  // small products stay on the CUDA-core kernel
  if (M < 64 || N < 32 || K < 32) return -1;
  static int sm_count = 0;
  static bool attr_set = false;
  if (sm_count == 0) {
    int dev = 0;
    AVSR_CHECK_CUDA(cudaGetDevice(&dev));
    AVSR_CHECK_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
  }
  if (!attr_set) {
    AVSR_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_set = true;
  }
  Params p;
  p.M = M; p.N = N; p.K = K;
  p.block_n = N > 128 ? 256 : (N > 64 ? 128 : (N > 32 ? 64 : 32));
  p.tiles_m = cdiv(M, BLOCK_M);
  p.k_tiles = cdiv(K, BLOCK_K);
  // few output tiles and a shallow K (per-step recurrent products): narrower N tiles spread the work over more
  // SMs; these launches are latency-bound, not tensor-pipe bound.  With a DEEP K (weight gradients: K = T*B) split-K
  // fills the SMs instead and the tiles stay wide: a 128 x 64 tile re-reads every A element 16 times and the launch is
  // bound by the 1.8 GB it pulls through L2, a 128 x 256 tile halves that.
  const bool deep = p.k_tiles >= 128 && !round_out;
  while (!deep && p.block_n > 64 && p.tiles_m * cdiv(N, p.block_n) < sm_count / 2) p.block_n >>= 1;
  p.tiles_n = cdiv(N, p.block_n);
  const int tiles = p.tiles_m * p.tiles_n;
  int splitk = 1;
  if (tiles < sm_count && p.k_tiles >= 16 && !round_out) {
    // as many K splits as fit in ONE wave of CTAs (a 149th work item would double the kernel time)
    splitk = min(p.k_tiles / 8, sm_count / tiles);
    if (splitk < 1) splitk = 1;
  }
  p.k_tiles_per_split = cdiv(p.k_tiles, splitk);
  p.splitk = cdiv(p.k_tiles, p.k_tiles_per_split);  // no empty splits
  p.num_work = tiles * p.splitk;
  p.a_mn = transA ? 1 : 0;
  p.b_mn = transB ? 0 : 1;
  p.reduce = (beta == 1.0f || p.splitk > 1) ? 1 : 0;
  p.bias = bias;
  p.round_out = (round_out && beta == 0.0f) ? 1 : 0;
  // instruction descriptor (cute::UMMA::InstrDescriptor): D=f32, A=B=tf32, majors, N>>3, M>>4
  p.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)p.a_mn << 15) | ((uint32_t)p.b_mn << 16) |
            ((uint32_t)(p.block_n >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
  CUtensorMap tmA, tmB, tmC;
  int bad = 0;
  if (!p.a_mn) bad |= make_map(&tmA, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, BLOCK_K, BLOCK_M);
  else bad |= make_map(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 32, BLOCK_K, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  if (!p.b_mn) bad |= make_map(&tmB, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, BLOCK_K, (uint32_t)p.block_n);
  else bad |= make_map(&tmB, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, 32, BLOCK_K, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B);
  bad |= make_map(&tmC, C, (uint64_t)N, (uint64_t)M, (uint64_t)ldc, 32, 32);
  if (bad) return -1;  // driver entry point unavailable or layout rejected
  if (p.splitk > 1 && beta == 0.0f) {
    // split-K partial sums are reduce-added: clear the destination block first
    AVSR_CHECK_CUDA(cudaMemset2DAsync(C, (size_t)ldc * sizeof(float), 0, (size_t)N * sizeof(float), (size_t)M, st));
  }
  const int grid = min(p.num_work, sm_count);
  const int slot = kernel_timer_begin(st, AVSR_K_GEMM);
  gemm_tc_kernel<<<grid, THREADS, SMEM_BYTES, st>>>(tmA, tmB, tmC, p);
  kernel_timer_end(st, slot);
  ++g_launch_count;
  AVSR_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace avsr
