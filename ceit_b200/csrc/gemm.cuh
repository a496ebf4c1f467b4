// Dense row-major GEMM used by every dense step of the path (block Cholesky, selected inverse,
// Schur complements, Y_i = Xh Sinv_i, inverse model, batched reconstruction):
//     C[b] = alpha * op(A[b]) * op(B[b]) + beta * C[b],   b = 0..batch-1 (strided)
//   * f64: FP64 tensor cores, mma.sync.aligned.m8n8k4.f64 (SASS DMMA).  tcgen05 has no f64
//     kind, so on sm_100a DMMA *is* the FP64 tensor path.
//   * complex f64: the same DMMA kernel structure on split re/im planes (four DMMA chains per complex
//     tile product); a SIMT FMA tile kernel is kept as a cross-check (force_simt_gemm).
#pragma once
#include <algorithm>
#include <atomic>
#include "scalar.cuh"

namespace ceit {

extern std::atomic<int64_t> g_launches;
extern int g_force_simt;
extern int g_gemm_tile;
extern int g_gemm_min_ctas64;   // 64x64 tiles when they give at least this many CTAs (else 32x32)
// Algorithmic real flops of the dense work enqueued so far (host-side count, single caller thread): 2 M N K per
// real product (8 M N K complex; half for a lower-only symmetric result), 2/3 n^3 per tile factorisation + inverse.
// ceit_stage_flops() reports the per-stage differences next to the stage timers (roofline numerators).
extern double g_flops;
extern int g_gemm_vec;   // 1 (default): 16-byte cp.async staging where the operand layout allows it

#define CEIT_COUNT_LAUNCH() (::ceit::g_launches.fetch_add(1, std::memory_order_relaxed))

// ------------------------------------------------------------------------------------------------
// FP64 tensor-core kernel, templated on the CTA tile (BM x BN, BK = 16), the warp grid (WM x WN) and
// the number of cp.async stages.  Each warp owns a (BM/WM) x (BN/WN) block of m8n8k4 accumulators.
// Operands are copied global -> shared with 8-byte cp.async (zero-fill out of bounds, so any M, N, K,
// lda work) in their natural layout; the transpose variants differ only in how fragments are read.
// Shared-memory row strides are == 4 (mod 16) doubles so the 16 lanes of a half-warp that fetch one
// fragment hit 16 distinct 8-byte banks.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma_884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc, bool valid) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int sz = valid ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(d), "l"(gsrc), "r"(sz));
}
// 16-byte copy of which the first `bytes` (0, 8 or 16) come from global memory, the rest is zero-filled
__device__ __forceinline__ void cp_async16_partial(double* smem_dst, const double* gsrc, int bytes) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "r"(bytes));
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gsrc), "r"(sz));
}
// one element of either scalar type (zero fill when !valid)
__device__ __forceinline__ void cp_async_elem(double* smem_dst, const double* gsrc, bool valid) {
  cp_async8(smem_dst, gsrc, valid);
}
__device__ __forceinline__ void cp_async_elem(cd* smem_dst, const cd* gsrc, bool valid) {
  cp_async16(smem_dst, gsrc, valid);
}
// ---- programmatic dependent launch (PDL) ---------------------------------------------------------------------------
// The hot path is a chain of ~100 short dependent kernels.  A kernel launched through launch_pdl() may be SCHEDULED
// while its same-stream predecessor still runs (the predecessor signals pdl_trigger() at its start, i.e. as soon as all
// of its CTAs are resident); it blocks in pdl_wait() until the predecessor has completed and its writes are visible.
// Everything before pdl_wait() - parameter reads, index arithmetic, shared-memory set-up - overlaps the predecessor's
// tail, and the launch latency itself is hidden.  Both instructions are no-ops for a normally launched kernel, and a
// kernel launched normally after a triggering one is serialised as usual, so kernels can adopt this one by one.
// ceit_set_option("pdl", 0) switches the launch attribute off, 1 (default) = GEMM + tile kernels, 2 = every kernel.
extern int g_pdl;
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;\n" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory"); }
template <class... KArgs, class... Args>
inline cudaError_t launch_pdl_level(int level, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                    cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = g_pdl >= level ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
// level 1 (default): the DMMA GEMM and the tile factorisation - the kernels that make up the serial spine
template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args... args) {
  return launch_pdl_level(1, kernel, grid, block, smem, st, args...);
}
// pdl(kernel, grid, block, smem, stream)(args...) - drop-in for the triple-chevron launch; the attribute is set from
// level 2 on (ceit_set_option("pdl", 2)): measured on MESH_50_50, programmatic edges between the short gather / scatter /
// element-wise kernels (whose first instruction already needs the predecessor's data) do not pay - step 0.675 ms with
// level 1, 0.685 ms with level 2, and a graph of 100 single-frame GEMVs 3.25 -> 3.68 us per frame
template <class... KArgs>
struct PdlLaunch {
  void (*k)(KArgs...);
  dim3 g, b;
  size_t s;
  cudaStream_t st;
  template <class... A>
  void operator()(A... a) const { launch_pdl_level(2, k, g, b, s, st, a...); }
};
template <class... KArgs>
inline PdlLaunch<KArgs...> pdl(void (*k)(KArgs...), dim3 g, dim3 b, size_t s, cudaStream_t st) {
  return PdlLaunch<KArgs...>{k, g, b, s, st};
}

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

template <int BM, int BN, int OPA, int OPB>
struct DmmaSmem {
  static constexpr int BK = 16;
  static constexpr int SA = (OPA == 0) ? BK + 4 : BM + 4;
  static constexpr int SB = (OPB == 0) ? BN + 4 : BK + 4;
  static constexpr int A_ELEMS = (OPA == 0) ? BM * SA : BK * SA;
  static constexpr int B_ELEMS = (OPB == 0) ? BK * SB : BN * SB;
  static constexpr int STAGE = A_ELEMS + B_ELEMS;
};

template <int BM, int BN, int WM, int WN, int OPA, int OPB, int STAGES>
__global__ void __launch_bounds__(WM * WN * 32)
gemm_dmma_kernel(int M, int N, int K, double alpha, const double* __restrict__ A, int64_t lda,
                 int64_t sA, const double* __restrict__ B, int64_t ldb, int64_t sB, double beta,
                 double* __restrict__ C, int64_t ldc, int64_t sC, int flags, int vec) {
  using SM = DmmaSmem<BM, BN, OPA, OPB>;
  pdl_trigger();
  // vec: bit 0 / bit 1 = the A / B operand is staged with 16-byte cp.async (two doubles along its contiguous
  // dimension; needs a 16-byte aligned base, an even leading dimension and an even batch stride - decided by the
  // host launcher); an 8-byte LDGSTS occupies the LSU as long as a 16-byte one, so this halves the staging issue
  // flags: bit 0 = lower_only; bits 1.. = inner dimension of the LAST batch entry (0 = K), so that the short last
  // slice of a split-K product runs in the same launch as the full ones
  if ((flags >> 1) != 0 && blockIdx.z == gridDim.z - 1) K = flags >> 1;
  const int lower_only = flags & 1;
  // vec bit 2: the grid is transposed (blockIdx.x walks the M tiles): consecutive CTAs then share one B tile and
  // sweep A, which streams the LARGER operand from HBM once when N >> M (batched reconstruction: dV is 192 MB, inv_mat
  // 7.8 MB - with the N tiles on the fast index every row of tiles re-read dV: 12.3 GB of DRAM reads for 0.2 GB of data)
  const unsigned tile_n = (vec & 4) ? blockIdx.y : blockIdx.x, tile_m = (vec & 4) ? blockIdx.x : blockIdx.y;
  // symmetric results whose consumers read the lower triangle only (A -= L L^T before a Cholesky):
  // tiles strictly above the block diagonal are skipped
  if (lower_only && (int)(tile_n * BN) > (int)(tile_m * BM + BM - 1)) return;
  constexpr int BK = 16, NT = WM * WN * 32;
  constexpr int SA = SM::SA, SB = SM::SB;
  constexpr int TM = BM / WM, TN = BN / WN, MI = TM / 8, NI = TN / 8;
  extern __shared__ __align__(16) double dsm[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int m0 = tile_m * BM, n0 = tile_n * BN;
  const int wm = (warp / WN) * TM, wn = (warp % WN) * TN;
  A += (int64_t)blockIdx.z * sA;
  B += (int64_t)blockIdx.z * sB;
  C += (int64_t)blockIdx.z * sC;

  double acc[MI][NI][2];
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int j = 0; j < NI; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  auto issue = [&](int kt, int buf) {
    double* As = dsm + buf * SM::STAGE;
    double* Bs = As + SM::A_ELEMS;
    const int k0 = kt * BK;
    if (vec & 1) {
#pragma unroll
      for (int r = 0; r < (BM * BK / 2) / NT; ++r) {
        const int idx = tid + r * NT;
        int m, k;                                   // first element of the pair along the contiguous dimension
        if (OPA == 0) {
          k = (idx & 7) * 2;
          m = idx >> 3;
        } else {
          m = (idx % (BM / 2)) * 2;
          k = idx / (BM / 2);
        }
        const int gm = m0 + m, gk = k0 + k;
        int cnt = 0;                                // valid elements of the pair
        if (gm < M && gk < K) cnt = (OPA == 0) ? (gk + 1 < K ? 2 : 1) : (gm + 1 < M ? 2 : 1);
        const double* src = cnt ? ((OPA == 0) ? A + (int64_t)gm * lda + gk : A + (int64_t)gk * lda + gm) : A;
        cp_async16_partial(As + ((OPA == 0) ? m * SA + k : k * SA + m), src, 8 * cnt);
      }
    } else {
#pragma unroll
      for (int r = 0; r < (BM * BK) / NT; ++r) {
        const int idx = tid + r * NT;
        int m, k;
        if (OPA == 0) {
          k = idx & 15;
          m = idx >> 4;
        } else {
          m = idx % BM;
          k = idx / BM;
        }
        const int gm = m0 + m, gk = k0 + k;
        const bool ok = (gm < M) && (gk < K);
        const double* src = ok ? ((OPA == 0) ? A + (int64_t)gm * lda + gk : A + (int64_t)gk * lda + gm) : A;
        cp_async8(As + ((OPA == 0) ? m * SA + k : k * SA + m), src, ok);
      }
    }
    if (vec & 2) {
#pragma unroll
      for (int r = 0; r < (BN * BK / 2) / NT; ++r) {
        const int idx = tid + r * NT;
        int n, k;
        if (OPB == 0) {
          n = (idx % (BN / 2)) * 2;
          k = idx / (BN / 2);
        } else {
          k = (idx & 7) * 2;
          n = idx >> 3;
        }
        const int gn = n0 + n, gk = k0 + k;
        int cnt = 0;
        if (gn < N && gk < K) cnt = (OPB == 0) ? (gn + 1 < N ? 2 : 1) : (gk + 1 < K ? 2 : 1);
        const double* src = cnt ? ((OPB == 0) ? B + (int64_t)gk * ldb + gn : B + (int64_t)gn * ldb + gk) : B;
        cp_async16_partial(Bs + ((OPB == 0) ? k * SB + n : n * SB + k), src, 8 * cnt);
      }
    } else {
#pragma unroll
      for (int r = 0; r < (BN * BK) / NT; ++r) {
        const int idx = tid + r * NT;
        int n, k;
        if (OPB == 0) {
          n = idx % BN;
          k = idx / BN;
        } else {
          k = idx & 15;
          n = idx >> 4;
        }
        const int gn = n0 + n, gk = k0 + k;
        const bool ok = (gn < N) && (gk < K);
        const double* src = ok ? ((OPB == 0) ? B + (int64_t)gk * ldb + gn : B + (int64_t)gn * ldb + gk) : B;
        cp_async8(Bs + ((OPB == 0) ? k * SB + n : n * SB + k), src, ok);
      }
    }
  };

  pdl_wait();                                       // operands come from the predecessor: block until it has completed
  const int nk = (K + BK - 1) / BK;
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < nk) issue(s, s);
    cp_async_commit();
  }
  for (int kt = 0; kt < nk; ++kt) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {
      const int nxt = kt + STAGES - 1;
      if (nxt < nk) issue(nxt, nxt % STAGES);
      cp_async_commit();
    }
    const double* As = dsm + (kt % STAGES) * SM::STAGE;
    const double* Bs = As + SM::A_ELEMS;
#pragma unroll
    for (int kk = 0; kk < BK; kk += 4) {
      double af[MI], bf[NI];
#pragma unroll
      for (int i = 0; i < MI; ++i) {
        const int m = wm + i * 8 + g, k = kk + t;
        af[i] = (OPA == 0) ? As[m * SA + k] : As[k * SA + m];
      }
#pragma unroll
      for (int j = 0; j < NI; ++j) {
        const int n = wn + j * 8 + g, k = kk + t;
        bf[j] = (OPB == 0) ? Bs[k * SB + n] : Bs[n * SB + k];
      }
#pragma unroll
      for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) dmma_884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
  }
  cp_async_wait<0>();
#pragma unroll
  for (int i = 0; i < MI; ++i) {
    const int row = m0 + wm + i * 8 + g;
    if (row >= M) continue;
#pragma unroll
    for (int j = 0; j < NI; ++j) {
      const int col = n0 + wn + j * 8 + 2 * t;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        if (col + q < N) {
          double* p = C + (int64_t)row * ldc + col + q;
          double v = alpha * acc[i][j][q];
          if (beta != 0.0) v = fma(beta, *p, v);
          *p = v;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Complex f64 on the FP64 tensor cores: the interleaved (re, im) operands are split into real and
// imaginary planes while they are copied to shared memory (two 8-byte cp.async per element), and every
// complex tile product is four DMMA chains:  Cr += Ar Br + (-Ai) Bi,  Ci += Ar Bi + Ai Br.
// Same tiling, pipeline and bank-conflict-free strides as the real kernel.
// ------------------------------------------------------------------------------------------------
template <int BM, int BN, int WM, int WN, int OPA, int OPB, int STAGES>
__global__ void __launch_bounds__(WM * WN * 32)
gemm_zdmma_kernel(int M, int N, int K, cd alpha, const cd* __restrict__ A, int64_t lda, int64_t sA,
                  const cd* __restrict__ B, int64_t ldb, int64_t sB, cd beta, cd* __restrict__ C, int64_t ldc,
                  int64_t sC, int beta_zero) {
  pdl_trigger();
  pdl_wait();
  using SM = DmmaSmem<BM, BN, OPA, OPB>;
  constexpr int BK = 16, NT = WM * WN * 32;
  constexpr int SA = SM::SA, SB = SM::SB;
  constexpr int TM = BM / WM, TN = BN / WN, MI = TM / 8, NI = TN / 8;
  constexpr int ZSTAGE = 2 * SM::STAGE;            // re + im planes of A and B
  extern __shared__ __align__(16) double dsm[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int wm = (warp / WN) * TM, wn = (warp % WN) * TN;
  A += (int64_t)blockIdx.z * sA;
  B += (int64_t)blockIdx.z * sB;
  C += (int64_t)blockIdx.z * sC;

  double cr[MI][NI][2], ci[MI][NI][2];
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int j = 0; j < NI; ++j) cr[i][j][0] = cr[i][j][1] = ci[i][j][0] = ci[i][j][1] = 0.0;

  auto issue = [&](int kt, int buf) {
    double* Ar = dsm + buf * ZSTAGE;
    double* Ai = Ar + SM::A_ELEMS;
    double* Br = Ai + SM::A_ELEMS;
    double* Bi = Br + SM::B_ELEMS;
    const int k0 = kt * BK;
#pragma unroll
    for (int r = 0; r < (BM * BK) / NT; ++r) {
      const int idx = tid + r * NT;
      int m, k;
      if (OPA == 0) {
        k = idx & 15;
        m = idx >> 4;
      } else {
        m = idx % BM;
        k = idx / BM;
      }
      const int gm = m0 + m, gk = k0 + k;
      const bool ok = (gm < M) && (gk < K);
      const double* src =
          reinterpret_cast<const double*>(ok ? ((OPA == 0) ? A + (int64_t)gm * lda + gk : A + (int64_t)gk * lda + gm) : A);
      const int so = (OPA == 0) ? m * SA + k : k * SA + m;
      cp_async8(Ar + so, src, ok);
      cp_async8(Ai + so, src + 1, ok);
    }
#pragma unroll
    for (int r = 0; r < (BN * BK) / NT; ++r) {
      const int idx = tid + r * NT;
      int n, k;
      if (OPB == 0) {
        n = idx % BN;
        k = idx / BN;
      } else {
        k = idx & 15;
        n = idx >> 4;
      }
      const int gn = n0 + n, gk = k0 + k;
      const bool ok = (gn < N) && (gk < K);
      const double* src =
          reinterpret_cast<const double*>(ok ? ((OPB == 0) ? B + (int64_t)gk * ldb + gn : B + (int64_t)gn * ldb + gk) : B);
      const int so = (OPB == 0) ? k * SB + n : n * SB + k;
      cp_async8(Br + so, src, ok);
      cp_async8(Bi + so, src + 1, ok);
    }
  };

  const int nk = (K + BK - 1) / BK;
#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < nk) issue(s, s);
    cp_async_commit();
  }
  for (int kt = 0; kt < nk; ++kt) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {
      const int nxt = kt + STAGES - 1;
      if (nxt < nk) issue(nxt, nxt % STAGES);
      cp_async_commit();
    }
    const double* Ar = dsm + (kt % STAGES) * ZSTAGE;
    const double* Ai = Ar + SM::A_ELEMS;
    const double* Br = Ai + SM::A_ELEMS;
    const double* Bi = Br + SM::B_ELEMS;
#pragma unroll
    for (int kk = 0; kk < BK; kk += 4) {
      double ar[MI], ai[MI], nai[MI], br[NI], bi[NI];
#pragma unroll
      for (int i = 0; i < MI; ++i) {
        const int m = wm + i * 8 + g, k = kk + t;
        const int so = (OPA == 0) ? m * SA + k : k * SA + m;
        ar[i] = Ar[so];
        ai[i] = Ai[so];
        nai[i] = -ai[i];
      }
#pragma unroll
      for (int j = 0; j < NI; ++j) {
        const int n = wn + j * 8 + g, k = kk + t;
        const int so = (OPB == 0) ? k * SB + n : n * SB + k;
        br[j] = Br[so];
        bi[j] = Bi[so];
      }
#pragma unroll
      for (int i = 0; i < MI; ++i)
#pragma unroll
        for (int j = 0; j < NI; ++j) {
          dmma_884(cr[i][j][0], cr[i][j][1], ar[i], br[j]);
          dmma_884(ci[i][j][0], ci[i][j][1], ar[i], bi[j]);
          dmma_884(cr[i][j][0], cr[i][j][1], nai[i], bi[j]);
          dmma_884(ci[i][j][0], ci[i][j][1], ai[i], br[j]);
        }
    }
  }
  cp_async_wait<0>();
#pragma unroll
  for (int i = 0; i < MI; ++i) {
    const int row = m0 + wm + i * 8 + g;
    if (row >= M) continue;
#pragma unroll
    for (int j = 0; j < NI; ++j) {
      const int col = n0 + wn + j * 8 + 2 * t;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        if (col + q < N) {
          cd* p = C + (int64_t)row * ldc + col + q;
          cd v = mul(alpha, mk(cr[i][j][q], ci[i][j][q]));
          if (!beta_zero) v = fma_(beta, *p, v);
          *p = v;
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Generic SIMT tile kernel (any scalar type): 64x64 tile, BK = 16, 256 threads, 4x4 per thread.
// ------------------------------------------------------------------------------------------------
template <class T, int OPA, int OPB>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(int M, int N, int K, T alpha, const T* __restrict__ A, int64_t lda, int64_t sA,
                 const T* __restrict__ B, int64_t ldb, int64_t sB, T beta, T* __restrict__ C,
                 int64_t ldc, int64_t sC, int beta_zero) {
  pdl_trigger();
  pdl_wait();
  constexpr int BM = 64, BN = 64, BK = 16;
  __shared__ T As[BK][BM + 1];
  __shared__ T Bs[BK][BN + 1];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  A += (int64_t)blockIdx.z * sA;
  B += (int64_t)blockIdx.z * sB;
  C += (int64_t)blockIdx.z * sC;
  T acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = Sc<T>::zero();
  for (int k0 = 0; k0 < K; k0 += BK) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      int m, k;
      if (OPA == 0) {
        k = tid & 15;
        m = (tid >> 4) + 16 * r;
      } else {
        m = tid & 63;
        k = (tid >> 6) + 4 * r;
      }
      int gm = m0 + m, gk = k0 + k;
      T v = Sc<T>::zero();
      if (gm < M && gk < K) v = (OPA == 0) ? A[(int64_t)gm * lda + gk] : A[(int64_t)gk * lda + gm];
      As[k][m] = v;
      int n;
      if (OPB == 0) {
        n = tid & 63;
        k = (tid >> 6) + 4 * r;
      } else {
        k = tid & 15;
        n = (tid >> 4) + 16 * r;
      }
      int gn = n0 + n;
      gk = k0 + k;
      v = Sc<T>::zero();
      if (gn < N && gk < K) v = (OPB == 0) ? B[(int64_t)gk * ldb + gn] : B[(int64_t)gn * ldb + gk];
      Bs[k][n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      T a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma_(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int row = m0 + ty * 4 + i;
    if (row >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int col = n0 + tx * 4 + j;
      if (col < N) {
        T* p = C + (int64_t)row * ldc + col;
        T v = mul(alpha, acc[i][j]);
        if (!beta_zero) v = fma_(beta, *p, v);
        *p = v;
      }
    }
  }
}

inline bool is_zero(double a) { return a == 0.0; }
inline bool is_zero(cd a) { return a.x == 0.0 && a.y == 0.0; }

template <class T>
inline void gemm_simt_launch(cudaStream_t st, int opA, int opB, int64_t M, int64_t N, int64_t K,
                             T alpha, const T* A, int64_t lda, int64_t sA, const T* B, int64_t ldb,
                             int64_t sB, T beta, T* C, int64_t ldc, int64_t sC, int batch) {
  dim3 grid((unsigned)((N + 63) / 64), (unsigned)((M + 63) / 64), (unsigned)batch), block(256);
  int bz = is_zero(beta) ? 1 : 0;
#define CEIT_SIMT_CASE(a, b)                                                                      \
  pdl(gemm_simt_kernel<T, a, b>, grid, block, 0, st)((int)M, (int)N, (int)K, alpha, A, lda, sA, B, \
                                                    ldb, sB, beta, C, ldc, sC, bz)
  if (opA == 0 && opB == 0) CEIT_SIMT_CASE(0, 0);
  else if (opA == 0 && opB == 1) CEIT_SIMT_CASE(0, 1);
  else if (opA == 1 && opB == 0) CEIT_SIMT_CASE(1, 0);
  else CEIT_SIMT_CASE(1, 1);
#undef CEIT_SIMT_CASE
  CEIT_COUNT_LAUNCH();
}

// C[b] = alpha op(A[b]) op(B[b]) + beta C[b]; strides sA/sB/sC in elements (0 = shared operand).
template <class T>
inline void gemm(cudaStream_t st, int opA, int opB, int64_t M, int64_t N, int64_t K, T alpha,
                 const T* A, int64_t lda, int64_t sA, const T* B, int64_t ldb, int64_t sB, T beta,
                 T* C, int64_t ldc, int64_t sC, int batch = 1, int lower_only = 0);   // specialised below

template <int BM, int BN, int WM, int WN, int STAGES>
inline void gemm_zdmma_launch(cudaStream_t st, int opA, int opB, int64_t M, int64_t N, int64_t K, cd alpha,
                              const cd* A, int64_t lda, int64_t sA, const cd* B, int64_t ldb, int64_t sB, cd beta,
                              cd* C, int64_t ldc, int64_t sC, int batch) {
  dim3 grid((unsigned)((N + BN - 1) / BN), (unsigned)((M + BM - 1) / BM), (unsigned)batch), block(WM * WN * 32);
  const int bz = is_zero(beta) ? 1 : 0;
#define CEIT_ZDMMA_CASE(a, b)                                                                              \
  {                                                                                                        \
    constexpr size_t smem = (size_t)STAGES * 2 * DmmaSmem<BM, BN, a, b>::STAGE * sizeof(double);           \
    static bool attr_done = false;                                                                         \
    if (!attr_done) {                                                                                      \
      cudaFuncSetAttribute(gemm_zdmma_kernel<BM, BN, WM, WN, a, b, STAGES>,                                \
                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                        \
      attr_done = true;                                                                                    \
    }                                                                                                      \
    pdl(gemm_zdmma_kernel<BM, BN, WM, WN, a, b, STAGES>, grid, block, smem, st)((int)M, (int)N, (int)K, alpha, A, lda, \
                                                                              sA, B, ldb, sB, beta, C, ldc, sC, bz); \
  }
  if (opA == 0 && opB == 0) CEIT_ZDMMA_CASE(0, 0)
  else if (opA == 0 && opB == 1) CEIT_ZDMMA_CASE(0, 1)
  else if (opA == 1 && opB == 0) CEIT_ZDMMA_CASE(1, 0)
  else CEIT_ZDMMA_CASE(1, 1)
#undef CEIT_ZDMMA_CASE
  CEIT_COUNT_LAUNCH();
}

template <>
inline void gemm<cd>(cudaStream_t st, int opA, int opB, int64_t M, int64_t N, int64_t K, cd alpha, const cd* A,
                     int64_t lda, int64_t sA, const cd* B, int64_t ldb, int64_t sB, cd beta, cd* C, int64_t ldc,
                     int64_t sC, int batch, int /*lower_only: full result for complex*/) {
  if (M <= 0 || N <= 0 || batch <= 0) return;
  g_flops += 8.0 * (double)M * (double)N * (double)K * batch;
  if (g_force_simt) {
    gemm_simt_launch<cd>(st, opA, opB, M, N, K, alpha, A, lda, sA, B, ldb, sB, beta, C, ldc, sC, batch);
    return;
  }
  auto ctas = [&](int64_t bm, int64_t bn) { return ((M + bm - 1) / bm) * ((N + bn - 1) / bn) * batch; };
  int cfg = g_gemm_tile;                       // 0 = auto, 1 = 32x32, 2/3 = 64x64
  if (cfg == 0) cfg = (ctas(64, 64) >= g_gemm_min_ctas64) ? 2 : 1;
  if (cfg >= 2)
    gemm_zdmma_launch<64, 64, 2, 2, 3>(st, opA, opB, M, N, K, alpha, A, lda, sA, B, ldb, sB, beta, C, ldc, sC, batch);
  else
    gemm_zdmma_launch<32, 32, 2, 2, 4>(st, opA, opB, M, N, K, alpha, A, lda, sA, B, ldb, sB, beta, C, ldc, sC, batch);
}

template <int BM, int BN, int WM, int WN, int STAGES>
inline void gemm_dmma_launch(cudaStream_t st, int opA, int opB, int64_t M, int64_t N, int64_t K, double alpha,
                             const double* A, int64_t lda, int64_t sA, const double* B, int64_t ldb, int64_t sB,
                             double beta, double* C, int64_t ldc, int64_t sC, int batch, int lower_only) {
  dim3 grid((unsigned)((N + BN - 1) / BN), (unsigned)((M + BM - 1) / BM), (unsigned)batch), block(WM * WN * 32);
  auto vec_ok = [](const double* p, int64_t ld, int64_t stride) {
    return ((reinterpret_cast<uintptr_t>(p) & 15) == 0) && (ld % 2 == 0) && (stride % 2 == 0);
  };
  int vec = g_gemm_vec ? ((vec_ok(A, lda, sA) ? 1 : 0) | (vec_ok(B, ldb, sB) ? 2 : 0)) : 0;
  if (N > M && grid.x <= 65535u && (double)K * (double)N * 8.0 > 32.0e6) {   // B is the big operand: M tiles fastest
    vec |= 4;
    grid = dim3(grid.y, grid.x, grid.z);
  }
#define CEIT_DMMA_CASE(a, b)                                                                              \
  {                                                                                                       \
    constexpr size_t smem = (size_t)STAGES * DmmaSmem<BM, BN, a, b>::STAGE * sizeof(double);              \
    static bool attr_done = false;                                                                        \
    if (!attr_done) {                                                                                     \
      cudaFuncSetAttribute(gemm_dmma_kernel<BM, BN, WM, WN, a, b, STAGES>,                                \
                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                       \
      attr_done = true;                                                                                   \
    }                                                                                                     \
    launch_pdl(gemm_dmma_kernel<BM, BN, WM, WN, a, b, STAGES>, grid, block, smem, st, (int)M, (int)N, (int)K, alpha, A,  \
               lda, sA, B, ldb, sB, beta, C, ldc, sC, lower_only, vec);                                               \
  }
  if (opA == 0 && opB == 0) CEIT_DMMA_CASE(0, 0)
  else if (opA == 0 && opB == 1) CEIT_DMMA_CASE(0, 1)
  else if (opA == 1 && opB == 0) CEIT_DMMA_CASE(1, 0)
  else CEIT_DMMA_CASE(1, 1)
#undef CEIT_DMMA_CASE
  CEIT_COUNT_LAUNCH();
}

// real products; cfg: 0 = automatic tile choice, 1 = 32x32, 2 = 64x64, 3 = 128x128.  `flags`: bit 0 = lower_only,
// bits 1.. = inner dimension of the last batch entry (see gemm_dmma_kernel)
inline void gemm_real_tiled(cudaStream_t st, int opA, int opB, int64_t M, int64_t N, int64_t K, double alpha,
                            const double* A, int64_t lda, int64_t sA, const double* B, int64_t ldb, int64_t sB,
                            double beta, double* C, int64_t ldc, int64_t sC, int batch, int flags, int cfg) {
  if (M <= 0 || N <= 0 || batch <= 0) return;
  {
    const double k_total = (flags >> 1) ? (double)(batch - 1) * K + (double)(flags >> 1) : (double)batch * K;
    g_flops += ((flags & 1) ? 1.0 : 2.0) * (double)M * (double)N * k_total;
  }
  if (g_force_simt) {
    gemm_simt_launch<double>(st, opA, opB, M, N, K, alpha, A, lda, sA, B, ldb, sB, beta, C, ldc, sC, batch);
    return;
  }
  auto ctas = [&](int64_t bm, int64_t bn) { return ((M + bm - 1) / bm) * ((N + bn - 1) / bn) * batch; };
  // measured on B200 (tools/time_gemm.py): the 64x64 tile (2 CTAs/SM) beats 128x128 (1 CTA/SM) at every size
  // of this path (30.3 vs 25.3 TFLOP/s on the inverse-model GEMM, cuBLAS 34.9); 32x32 when 64x64 cannot fill the chip
  if (cfg == 0) cfg = (ctas(64, 64) >= g_gemm_min_ctas64) ? 2 : 1;
  // (128 x 64 and 64 x 128 tiles with 8 warps, 2 CTAs/SM, were measured too: 25-29 TFLOP/s against 30-32 for the
  //  64 x 64 tile on the four large products of the path, tools/time_gemm_tiles.py)
  if (cfg == 3)
    gemm_dmma_launch<128, 128, 2, 4, 3>(st, opA, opB, M, N, K, alpha, A, lda, sA, B, ldb, sB, beta, C, ldc, sC, batch,
                                        flags);
  else if (cfg == 2)
    gemm_dmma_launch<64, 64, 2, 2, 3>(st, opA, opB, M, N, K, alpha, A, lda, sA, B, ldb, sB, beta, C, ldc, sC, batch,
                                      flags);   // 3 stages = 60 KB: 3 CTAs/SM
  else
    gemm_dmma_launch<32, 32, 2, 2, 4>(st, opA, opB, M, N, K, alpha, A, lda, sA, B, ldb, sB, beta, C, ldc, sC, batch,
                                      flags);
}

template <>
inline void gemm<double>(cudaStream_t st, int opA, int opB, int64_t M, int64_t N, int64_t K,
                         double alpha, const double* A, int64_t lda, int64_t sA, const double* B,
                         int64_t ldb, int64_t sB, double beta, double* C, int64_t ldc, int64_t sC,
                         int batch, int lower_only) {
  gemm_real_tiled(st, opA, opB, M, N, K, alpha, A, lda, sA, B, ldb, sB, beta, C, ldc, sC, batch, lower_only & 1,
                  g_gemm_tile);           // g_gemm_tile: 0 = auto unless forced through ceit_set_option("gemm_tile")
}


// ------------------------------------------------------------------------------------------------
// Split-K for small outputs with a long inner dimension (S_U = K_UU - K_0U^T X, G = Jt Jt^T): the K range
// is cut into `S` slices that run as one strided batch into partial results, then summed in fixed order.
// ------------------------------------------------------------------------------------------------
__global__ void splitk_reduce_kernel(int64_t n, int S, const double* __restrict__ P, double beta, double* __restrict__ C,
                                     int64_t M, int64_t N, int64_t ldc, int lower_only) {
  pdl_trigger();
  pdl_wait();
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  const int64_t r = q / N, c = q - r * N;
  if (lower_only && c > r) return;                    // the slices only hold the tiles on and below the diagonal
  double acc = 0.0;
  for (int s = 0; s < S; ++s) acc += P[(int64_t)s * n + q];
  double* p = C + r * ldc + c;
  *p = (beta != 0.0) ? fma(beta, *p, acc) : acc;
}
// A[r][c] = A[c][r] for c > r (after a lower-only symmetric product)
__global__ void symmetrize_from_lower_kernel(int64_t n, double* __restrict__ A, int64_t lda) {
  pdl_trigger();
  pdl_wait();
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n * n) return;
  const int64_t r = q / n, c = q - r * n;
  if (c > r) A[r * lda + c] = A[c * lda + r];
}

// Number of K slices (not monotone in K: a workspace shared by several inner dimensions is sized with
// splitk_max_slices).
//  * small outputs (< 384 wide): 32x32 tiles, ~4 CTAs per SM, slices >= 256 deep;
//  * large outputs whose 64x64 tiles (only those on/below the diagonal when lower_only) cannot fill the chip
//    (the 400x400-mesh Gram matrix: 136 tiles with K = 257 762; its S_U: 45 tiles with K = 160 000):
//    slices >= 2048 deep, ~4 CTAs per SM.
inline bool splitk_large(int64_t M, int64_t N) { return (M < N ? M : N) >= 384; }
inline int splitk_slices(int64_t M, int64_t N, int64_t K, int lower_only = 0) {
  if (splitk_large(M, N)) {
    const int64_t tm = (M + 63) / 64, tn = (N + 63) / 64;
    const int64_t t64 = (lower_only && M == N) ? tm * (tm + 1) / 2 : tm * tn;
    if (t64 >= 444 || K < 8192) return 1;
    // 444 CTAs are resident (148 SMs x 3): pick the slice count whose t64 * S CTAs leave the smallest idle tail in
    // their last wave (the 400x400 Gram matrix: 136 tiles x 5 slices = 1.53 waves ran at 77 %; x 13 = 3.98 waves)
    const int64_t maxS = std::min<int64_t>(K / 2048, 32);
    int64_t best = 1;
    double best_eff = 0.0;
    for (int64_t S = 2; S <= maxS; ++S) {
      const int64_t ctas = t64 * S, waves = (ctas + 443) / 444;
      const double eff = (double)ctas / (double)(waves * 444);
      if (eff > best_eff + 0.02) { best_eff = eff; best = S; }      // prefer fewer slices unless clearly better
    }
    return (int)best;
  }
  const int64_t tiles = ((M + 31) / 32) * ((N + 31) / 32);
  if (K < 1024 || tiles >= 296) return 1;
  int64_t S = (592 + tiles - 1) / tiles;              // aim at ~4 CTAs per SM
  const int64_t maxS = K / 256;                       // keep slices >= 256 deep
  if (S > maxS) S = maxS;
  if (S > 64) S = 64;
  return (int)(S < 2 ? 1 : S);
}

// upper bound of splitk_slices over every inner dimension <= Kmax
inline int splitk_max_slices(int64_t M, int64_t N, int64_t Kmax) {
  const int64_t cap = splitk_large(M, N) ? std::min<int64_t>(Kmax / 2048, 32) : std::min<int64_t>(Kmax / 256, 64);
  return (int)std::max<int64_t>(cap, 1);
}

// workspace: at least splitk_slices(M, N, K, lower_only) * M * N doubles (ignored when no split is used).
// lower_only (M == N, symmetric result): only the entries on and below the diagonal of C are produced.
inline void gemm_splitk(cudaStream_t st, int opA, int opB, int64_t M, int64_t N, int64_t K, double alpha, const double* A,
                        int64_t lda, const double* B, int64_t ldb, double beta, double* C, int64_t ldc,
                        double* workspace, int lower_only = 0) {
  const int S = workspace ? splitk_slices(M, N, K, lower_only) : 1;
  if (S <= 1) {
    gemm<double>(st, opA, opB, M, N, K, alpha, A, lda, 0, B, ldb, 0, beta, C, ldc, 0, 1, lower_only);
    return;
  }
  int64_t Ks = ((K + S - 1) / S + 15) / 16 * 16;
  const int full = (int)(K / Ks);                     // slices of exactly Ks
  const int64_t rem = K - (int64_t)full * Ks;
  const int64_t sA = (opA == 0) ? Ks : Ks * lda, sB = (opB == 0) ? Ks * ldb : Ks;
  // every slice on 64x64 tiles when the output is large (the slice count was sized for them)
  const int cfg = (splitk_large(M, N) && g_gemm_tile == 0) ? 2 : g_gemm_tile;
  // one launch: `full` slices of Ks and, when K is not a multiple, a last slice of `rem` (kernel flags bits 1..)
  int parts = full;
  if (rem > 0) ++parts;
  if (!g_force_simt) {
    gemm_real_tiled(st, opA, opB, M, N, Ks, alpha, A, lda, sA, B, ldb, sB, 0.0, workspace, N, M * N, parts,
                    (lower_only & 1) | (rem > 0 ? (int)(rem << 1) : 0), cfg);
  } else {                                            // the SIMT cross-check kernel has no tail slice
    if (full > 0) gemm<double>(st, opA, opB, M, N, Ks, alpha, A, lda, sA, B, ldb, sB, 0.0, workspace, N, M * N, full);
    if (rem > 0)
      gemm<double>(st, opA, opB, M, N, rem, alpha, A + (int64_t)full * sA, lda, 0, B + (int64_t)full * sB, ldb, 0, 0.0,
                   workspace + (int64_t)full * M * N, N, 0, 1);
  }
  pdl(splitk_reduce_kernel, (unsigned)((M * N + 255) / 256), 256, 0, st)(M * N, parts, workspace, beta, C, M, N, ldc,
                                                                       lower_only);
  CEIT_COUNT_LAUNCH();
}

}  // namespace ceit
