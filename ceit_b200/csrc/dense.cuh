// Dense symmetric factorisation primitives: Cholesky L L^T = A (plain transpose; with the complex
// square root this also covers the complex-SYMMETRIC matrices of the non-zero-base path) together
// with the explicit triangular inverse Linv = L^-1, so that every triangular solve of the path
// becomes a tensor-core GEMM.  Batched over blockIdx.(x) with strides.
#pragma once
#include "gemm.cuh"

namespace ceit {

constexpr int TILE = 64;
// phase clocks of the tile kernel (tools/tile_clocks.py); compiled in only with -DCEIT_TILE_CLOCKS
__device__ long long g_tile_clk[8];
#ifdef CEIT_TILE_CLOCKS
#define CEIT_CLK(k) do { if (threadIdx.x == 0 && blockIdx.x == 0) g_tile_clk[k] = clock64(); } while (0)
#else
#define CEIT_CLK(k) do { } while (0)
#endif

// Two pivot columns per barrier: phase PH (columns 16 PH .. 16 PH + 15) of the unscaled elimination of a
// 64 x 64 tile, done as 8 eliminations of a 2 x 2 pivot block P = [[p00, p10], [p10, p11]] (columns j, j+1):
//   C -= [a0 a1] P^-1 [b0 b1]^T,  P^-1 = [[p11, -p10], [-p10, p00]] / det.
// Thread (rb, cb) holds c[u][v] = S[rb + 16 u][cb + 16 v]; only the blocks u >= v, v >= PH are live (lower
// triangle, columns not yet eliminated).  Shared memory holds columns j and j+1 updated through column j-1
// (published by their owners at the end of the previous double step); every thread forms the multipliers
// itself, so a double step has ONE barrier and ONE reciprocal on its dependent chain.  The second column of a
// pair is frozen in its published state; the rank-1 update by its partner that it still lacks is applied
// after the last phase (tile_fix_odd_columns).
template <class T, int PH>
__device__ __forceinline__ void tile_eliminate_phase2(T* __restrict__ S, T (&c)[4][4], int rb, int cb, int n) {
  constexpr int LD = TILE + 1;
  constexpr int PN = PH < 3 ? PH + 1 : 3;
#pragma unroll 1
  for (int jj = 0; jj < 16; jj += 2) {
    const int j = 16 * PH + jj;
    if (j >= n) break;                       // n odd: the identity padding makes the extra column a no-op
    const T p00 = S[j * LD + j], p10 = S[(j + 1) * LD + j], p11 = S[(j + 1) * LD + j + 1];
    T a0[4], a1[4], t0[4], t1[4];
#pragma unroll
    for (int u = PH; u < 4; ++u) {
      a0[u] = S[(rb + 16 * u) * LD + j];
      a1[u] = S[(rb + 16 * u) * LD + j + 1];
    }
#pragma unroll
    for (int v = PH; v < 4; ++v) {
      t0[v] = S[(cb + 16 * v) * LD + j];
      t1[v] = S[(cb + 16 * v) * LD + j + 1];
    }
    const T rdet = rcp_fast(sub(mul(p00, p11), mul(p10, p10)));
    const T q11 = mul(p11, rdet), q10 = mul(p10, rdet), q00 = mul(p00, rdet);
#pragma unroll
    for (int v = PH; v < 4; ++v) {
      const T b0 = t0[v], b1 = t1[v];
      t0[v] = sub(mul(q11, b0), mul(q10, b1));
      t1[v] = sub(mul(q00, b1), mul(q10, b0));
    }
    if (cb <= jj + 1) {                      // my column of this phase is final (or frozen: second of the pair)
      t0[PH] = Sc<T>::zero();
      t1[PH] = Sc<T>::zero();
    }
#pragma unroll
    for (int v = PH; v < 4; ++v)
#pragma unroll
      for (int u = v; u < 4; ++u) c[u][v] = sub(sub(c[u][v], mul(a0[u], t0[v])), mul(a1[u], t1[v]));
    if (jj < 14) {
      if (cb == jj + 2 || cb == jj + 3) {
#pragma unroll
        for (int u = PH; u < 4; ++u) S[(rb + 16 * u) * LD + 16 * PH + cb] = c[u][PH];
      }
    } else if (PH < 3) {
      if (cb < 2) {
#pragma unroll
        for (int u = PN; u < 4; ++u) S[(rb + 16 * u) * LD + 16 * PN + cb] = c[u][PN];
      }
    }
    __syncthreads();
  }
}
// Odd columns k were frozen one rank-1 update short: c[.][k] -= S[.][k-1] S[k][k-1] / S[k-1][k-1] (column k-1 is
// final in shared memory).
template <class T>
__device__ __forceinline__ void tile_fix_odd_columns(const T* __restrict__ S, T (&c)[4][4], int rb, int cb) {
  constexpr int LD = TILE + 1;
  if (cb & 1) {
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int k = cb + 16 * v;
      const T m = mul(S[k * LD + k - 1], rcp_fast(S[(k - 1) * LD + k - 1]));
#pragma unroll
      for (int u = v; u < 4; ++u) c[u][v] = sub(c[u][v], mul(S[(rb + 16 * u) * LD + k - 1], m));
    }
  }
}

// One level of the block recursion for the inverse of a lower-triangular 64x64 tile held in shared
// memory (leading dimension 65): for every pair of diagonal blocks of size HALF,
//   W = L21 X11 (X11 lower triangular),  X21 = -X22 W (X22 lower triangular).
// 256 threads; each thread produces NOUT = 64*HALF/256 outputs of one column so that the k-loop carries
// NOUT independent accumulators.
template <class T, int HALF>
__device__ __forceinline__ void tri_inverse_level(const T* __restrict__ S, T* __restrict__ X, T* __restrict__ W,
                                                  int tid) {
  constexpr int LD = TILE + 1;
  constexpr int NPAIR = TILE / (2 * HALF);
  constexpr int NOUT = (NPAIR * HALF * HALF) / 256;   // 2 for HALF=16, 4 for HALF=32
  constexpr int RSTEP = HALF / NOUT;
  const int c = tid % HALF;
  const int rest = tid / HALF;            // 0 .. 256/HALF-1
  const int r0 = rest % RSTEP;            // base row inside the block
  const int pr = rest / RSTEP;            // pair index
  const int o = pr * 2 * HALF;
  T acc[NOUT];
#pragma unroll
  for (int u = 0; u < NOUT; ++u) acc[u] = Sc<T>::zero();
#pragma unroll 8
  for (int k = 0; k < HALF; ++k) {                 // X11 is lower triangular: rows k < c are exact zeros
    const T xv = X[(o + k) * LD + o + c];
#pragma unroll
    for (int u = 0; u < NOUT; ++u) acc[u] = fma_(S[(o + HALF + r0 + u * RSTEP) * LD + o + k], xv, acc[u]);
  }
#pragma unroll
  for (int u = 0; u < NOUT; ++u) W[(o + HALF + r0 + u * RSTEP) * LD + o + c] = acc[u];
  __syncthreads();
#pragma unroll
  for (int u = 0; u < NOUT; ++u) acc[u] = Sc<T>::zero();
#pragma unroll 8
  for (int k = 0; k < HALF; ++k) {
    const T wv = W[(o + HALF + k) * LD + o + c];
#pragma unroll
    for (int u = 0; u < NOUT; ++u)   // X22 is lower triangular: entries with k > row are exact zeros
      acc[u] = fma_(X[(o + HALF + r0 + u * RSTEP) * LD + o + HALF + k], wv, acc[u]);
  }
#pragma unroll
  for (int u = 0; u < NOUT; ++u) X[(o + HALF + r0 + u * RSTEP) * LD + o + c] = neg(acc[u]);
  __syncthreads();
}

// Factorisation + triangular inverse of a 64 x 64 tile held in shared memory (leading dimension 65).
// In: S = the matrix (identity-padded beyond n), X = 0, all threads synchronised.  Out: S = L (zeros above the
// diagonal), X = L^-1, synchronised.  W: scratch tile; rdiag / rpiv: 64 / 65 scalars.  256 threads.
template <class T>
__device__ __forceinline__ void tile_factor_in_smem(T* __restrict__ S, T* __restrict__ X, T* __restrict__ W,
                                                    T* __restrict__ rdiag, T* __restrict__ rpiv, int n, int tid,
                                                    int* __restrict__ info) {
  constexpr int LD = TILE + 1;
  // ---- unscaled right-looking elimination, one barrier per TWO columns -----------------------------
  // Thread (rb, cb) owns the 4x4 interleaved register tile {rb+16u} x {cb+16v} for the whole
  // elimination (interleaving balances the triangular work and keeps shared accesses conflict-free).
  // Per double step it loads 2 x 4 values of the two pivot columns for its rows, 2 x 4 for its columns and
  // the 2 x 2 pivot block, for 32 FMAs.  A column pair is published to shared memory once, when the
  // eliminations before it are complete.  Entries above the diagonal are carried along unpredicated: they
  // only feed other above-diagonal entries, which are never read for results.
  const int cb = tid & 15, rb = tid >> 4;
  T c[4][4];
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) c[u][v] = S[(rb + 16 * u) * LD + cb + 16 * v];
  __syncthreads();
  // The 64 steps are split into 4 phases of 16 (PH = j >> 4, a compile-time constant inside a phase), so
  // that "which of my columns are still live" and "which register holds the next pivot column" are static:
  // columns v < PH are final, v > PH always live, v == PH live iff cb > (j & 15).
  tile_eliminate_phase2<T, 0>(S, c, rb, cb, n);
  tile_eliminate_phase2<T, 1>(S, c, rb, cb, n);
  tile_eliminate_phase2<T, 2>(S, c, rb, cb, n);
  tile_eliminate_phase2<T, 3>(S, c, rb, cb, n);
  tile_fix_odd_columns<T>(S, c, rb, cb);
  CEIT_CLK(2);
  // final pivots: the diagonal entry (cb + 16 v, cb + 16 v) lives in c[v][v] of the thread with rb == cb
  if (rb == cb) {
#pragma unroll
    for (int v = 0; v < 4; ++v) rpiv[cb + 16 * v] = c[v][v];
  }
  __syncthreads();
  // ---- L = S diag(p)^-1/2 ; rdiag = 1 / L_jj --------------------------------------------------------
  {
    T rs[4];
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int k = cb + 16 * v;
      const T p = rpiv[k];
      if (rb == 0 && bad_pivot(p)) atomicExch(info, 1);
      rs[v] = inv_(sqrt_(p));
      if (rb == 0) rdiag[k] = rs[v];
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int v = 0; v < 4; ++v)
        S[(rb + 16 * u) * LD + cb + 16 * v] = (rb + 16 * u >= cb + 16 * v) ? mul(c[u][v], rs[v]) : Sc<T>::zero();
  }
  __syncthreads();
  CEIT_CLK(3);
  // ---- X = L^-1: four 16x16 diagonal blocks by forward substitution, column in registers ---------------
  if (tid < 64) {
    const int o = (tid >> 4) << 4, cc = tid & 15;
    // column-oriented (axpy) form: x_r is final after ONE dependent multiply, then updates all later rows
    // with independent FMAs, so the dependent chain is 2 operations per row instead of r
    T s[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) s[r] = (r == cc) ? Sc<T>::one() : Sc<T>::zero();
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const T x = mul(s[r], rdiag[o + r]);
      X[(o + r) * LD + o + cc] = x;
#pragma unroll
      for (int r2 = r + 1; r2 < 16; ++r2) s[r2] = sub(s[r2], mul(S[(o + r2) * LD + o + r], x));
    }
  }
  __syncthreads();
  // ---- recursion X21 = -X22 (L21 X11): half = 16 (two pairs, 2 outputs/thread), half = 32 (4/thread) ----
  CEIT_CLK(4);
  // (a tile of n <= 16 / n <= 32 is identity beyond its first / first two diagonal blocks: nothing to recurse on)
  if (n > 16) tri_inverse_level<T, 16>(S, X, W, tid);
  if (n > 32) tri_inverse_level<T, 32>(S, X, W, tid);
}

// One CTA factors one n x n (n <= 64) tile held in shared memory.
//   A (read), L (write, lower incl. zeros above the diagonal), Linv (write, same shape).
// info: set to 1 if a pivot is not positive (real) / zero (complex).
//
// Elimination is done UNSCALED (S[i][k] -= S[i][j] S[k][j] / S[j][j]) so that a column step needs one
// barrier and no serial sqrt: the pivots are final when their column is reached, and L = S diag(p)^-1/2
// is applied in one pass at the end.  The triangular inverse is built by block recursion
// (16 -> 32 -> 64): X21 = -X22 (L21 X11), i.e. small matrix products that use all 256 threads.
template <class T>
__global__ void __launch_bounds__(256)
potrf_trtri_tile_kernel(int n, const T* __restrict__ A, int64_t lda, int64_t sA, T* __restrict__ L,
                        int64_t ldl, int64_t sL, T* __restrict__ Linv, int64_t ldi, int64_t sI,
                        int* __restrict__ info) {
  pdl_trigger();
  pdl_wait();
  constexpr int LD = TILE + 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* S = reinterpret_cast<T*>(smem_raw);   // [64][65] matrix -> L
  T* X = S + TILE * LD;                    // [64][65] inverse
  T* W = X + TILE * LD;                    // [64][65] scratch for the recursion
  __shared__ T rdiag[TILE];
  __shared__ T rpiv[TILE + 1];
  const int tid = threadIdx.x;
  A += (int64_t)blockIdx.x * sA;
  L += (int64_t)blockIdx.x * sL;
  Linv += (int64_t)blockIdx.x * sI;
  CEIT_CLK(0);
  {
    T v[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) {   // all 16 global loads in flight before the first shared store
      const int r = (tid >> 6) + 4 * u, cc = tid & 63;
      v[u] = (r == cc) ? Sc<T>::one() : Sc<T>::zero();   // identity padding for n < 64
      if (r < n && cc < n) v[u] = A[(int64_t)r * lda + cc];
    }
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int r = (tid >> 6) + 4 * u, cc = tid & 63;
      S[r * LD + cc] = v[u];
      X[r * LD + cc] = Sc<T>::zero();
    }
  }
  __syncthreads();
  CEIT_CLK(1);
  tile_factor_in_smem<T>(S, X, W, rdiag, rpiv, n, tid, info);
  CEIT_CLK(5);
  // the caller zero-initialises L and Linv, so only the lower triangle is written
  for (int q = tid; q < TILE * TILE; q += 256) {
    const int r = q >> 6, cc = q & 63;
    if (r < n && cc <= r) {
      L[(int64_t)r * ldl + cc] = S[r * LD + cc];
      Linv[(int64_t)r * ldi + cc] = X[r * LD + cc];
    }
  }
  CEIT_CLK(6);
}


// ---- multi-level nested dissection: the whole front step of a level of SMALL fronts in one kernel -------------------
// (factor_tree.h, levels with pull = 1: padded own size s <= 64, padded boundary size b <= 64.)  One CTA per front:
//   1. load the own block A (s x s, identity-padded to 64) and the coupling block Bt (s x b) from the assembled
//      matrix, C = 0, then PULL the update matrices of the children (in the child order of the plan; the entries of
//      one child hit distinct destinations, a barrier separates the children: no atomics, bitwise reproducible);
//   2. A = L L^T and Linv in shared memory (tile_factor_in_smem);
//   3. A^-1 = Linv^T Linv -> global;  Y = Linv Bt;  W = A^-1 Bt -> global;  U = C - Y^T Y -> global C (lower triangle,
//      what the parent pulls / the extend-add pushes);  the updated Bt -> global (right-hand-side chain).
// Replaces 6-8 dependent launches per level (tile factorisation, 4 GEMMs, 2 extend-add passes) on the serial spine
// of the base stage: MESH_50_50 has 7 levels, all of this kind.
struct FrontFusedArgs {
  int s, b, f0;                 // padded sizes of the level, global id of its first front
  const int32_t *child_ptr, *child_ids, *f_b, *rel;
  const int64_t *f_offC, *f_relOff;
  const int32_t* inv_rel;       // destination-centric pull: parent front index -> boundary slot of the child
  const int64_t* f_invOff;
};
constexpr int FRONT_LD = TILE + 1;
constexpr size_t FRONT_SMEM = (size_t)(6 * TILE * FRONT_LD + 2 * TILE + 8) * sizeof(double);
__global__ void __launch_bounds__(256)
front_fused_kernel(FrontFusedArgs a, double* __restrict__ A, double* __restrict__ Bt, double* __restrict__ Call,
                   double* __restrict__ Cout, double* __restrict__ Ainv, double* __restrict__ Wg,
                   int* __restrict__ info) {
  pdl_trigger();
  pdl_wait();
  constexpr int LD = FRONT_LD;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* S = reinterpret_cast<double*>(smem_raw);   // matrix -> L
  double* X = S + TILE * LD;                          // Linv
  double* Wsc = X + TILE * LD;                        // scratch of the inverse recursion, then A^-1
  double* Bs = Wsc + TILE * LD;                       // Bt (s x b)
  double* Ys = Bs + TILE * LD;                        // Y = Linv Bt
  double* Cs = Ys + TILE * LD;                        // C (b x b)
  double* rdiag = Cs + TILE * LD;
  double* rpiv = rdiag + TILE;
  const int tid = threadIdx.x, f = blockIdx.x, s = a.s, b = a.b;
  const int g = a.f0 + f;
  A += (int64_t)f * s * s;
  Bt += (int64_t)f * s * b;
  Cout += (int64_t)f * b * b;
  Ainv += (int64_t)f * s * s;
  Wg += (int64_t)f * s * b;
  {
    double va[16], vb[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) {   // all global loads in flight before the first shared store
      const int r = (tid >> 6) + 4 * u, c = tid & 63;
      va[u] = (r == c) ? 1.0 : 0.0;
      if (r < s && c < s) va[u] = A[(int64_t)r * s + c];
      vb[u] = (r < s && c < b) ? Bt[(int64_t)r * b + c] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int r = (tid >> 6) + 4 * u, c = tid & 63;
      S[r * LD + c] = va[u];
      Bs[r * LD + c] = vb[u];
      X[r * LD + c] = 0.0;
      Cs[r * LD + c] = 0.0;
    }
  }
  __syncthreads();
  for (int qc = a.child_ptr[g]; qc < a.child_ptr[g + 1]; ++qc) {
    const int ch = a.child_ids[qc];
    const int cb = a.f_b[ch];                       // <= 64 is not guaranteed for a child: handled in chunks below
    const double* __restrict__ U = Call + a.f_offC[ch];
    const int32_t* __restrict__ rel = a.rel + a.f_relOff[ch];
    const int total = cb * cb;
    for (int base = 0; base < total; base += 256 * 8) {
      double v[8];
      int ra[8], rc[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {                 // 8 independent (rel, rel, U) load triples per thread in flight
        const int q = base + j * 256 + tid;
        ra[j] = rc[j] = -1;
        v[j] = 0.0;
        if (q < total) {
          const int x = q / cb, y = q - x * cb;
          ra[j] = rel[x];
          rc[j] = rel[y];
          v[j] = (x >= y) ? U[(int64_t)x * cb + y] : U[(int64_t)y * cb + x];
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (ra[j] < 0 || rc[j] < 0) continue;
        if (ra[j] < s) {
          if (rc[j] < s) S[ra[j] * LD + rc[j]] += v[j];
          else Bs[ra[j] * LD + (rc[j] - s)] += v[j];
        } else if (rc[j] >= s) {
          Cs[(ra[j] - s) * LD + (rc[j] - s)] += v[j];
        }
      }
    }
    __syncthreads();
  }
  tile_factor_in_smem<double>(S, X, Wsc, rdiag, rpiv, s, tid, info);
  // The four small products below use the register tiling of the elimination: thread (tr, tc) owns the 4 x 4
  // interleaved outputs {tr + 16 u} x {tc + 16 v}, i.e. 16 independent FMA chains per thread over the 64-deep inner
  // dimension (a single chain per output ran at the FMA latency: 3x slower than the tile factorisation itself).
  // Rows / columns beyond s or b are padding (identity / zero) and are masked at the stores.
  const int tr = tid >> 4, tc = tid & 15;
  double acc[4][4];
  // A^-1 = X^T X -> Wsc and global
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) acc[u][v] = 0.0;
  // (inner dimension: rows k >= s of X are the identity padding and contribute nothing to outputs r, c < s; the
  //  16 x 16 output blocks with u >= nu or v >= nu lie entirely in the padding and are skipped)
  const int nu = (s + 15) >> 4, nvb = (b + 15) >> 4;
#pragma unroll 4
  for (int k = 0; k < s; ++k) {
    double xr[4], xc[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) xr[u] = X[k * LD + tr + 16 * u];
#pragma unroll
    for (int v = 0; v < 4; ++v) xc[v] = X[k * LD + tc + 16 * v];
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (u < nu) {
#pragma unroll
        for (int v = 0; v < 4; ++v)
          if (v < nu) acc[u][v] = fma(xr[u], xc[v], acc[u][v]);
      }
  }
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int r = tr + 16 * u, c = tc + 16 * v;
      Wsc[r * LD + c] = (r < s && c < s) ? acc[u][v] : 0.0;
      if (r < s && c < s) Ainv[(int64_t)r * s + c] = acc[u][v];
    }
  if (b > 0) {
    // Y = X Bt
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int v = 0; v < 4; ++v) acc[u][v] = 0.0;
#pragma unroll 4
    for (int k = 0; k < s; ++k) {
      double xr[4], bc[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) xr[u] = X[(tr + 16 * u) * LD + k];
#pragma unroll
      for (int v = 0; v < 4; ++v) bc[v] = Bs[k * LD + tc + 16 * v];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (u < nu) {
#pragma unroll
          for (int v = 0; v < 4; ++v)
            if (v < nvb) acc[u][v] = fma(xr[u], bc[v], acc[u][v]);
        }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const int r = tr + 16 * u, c = tc + 16 * v;
        Ys[r * LD + c] = acc[u][v];
        if (r < s && c < b) Bt[(int64_t)r * b + c] = Bs[r * LD + c];   // coupling block incl. the children's part
      }
  }
  __syncthreads();
  if (b > 0) {
    // W = A^-1 Bt
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int v = 0; v < 4; ++v) acc[u][v] = 0.0;
#pragma unroll 4
    for (int k = 0; k < s; ++k) {
      double ar[4], bc[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) ar[u] = Wsc[(tr + 16 * u) * LD + k];
#pragma unroll
      for (int v = 0; v < 4; ++v) bc[v] = Bs[k * LD + tc + 16 * v];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (u < nu) {
#pragma unroll
          for (int v = 0; v < 4; ++v)
            if (v < nvb) acc[u][v] = fma(ar[u], bc[v], acc[u][v]);
        }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const int r = tr + 16 * u, c = tc + 16 * v;
        if (r < s && c < b) Wg[(int64_t)r * b + c] = acc[u][v];
      }
    // U = C - Y^T Y, lower triangle
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int v = 0; v < 4; ++v) acc[u][v] = Cs[(tr + 16 * u) * LD + tc + 16 * v];
#pragma unroll 4
    for (int k = 0; k < s; ++k) {
      double yx[4], yy[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) yx[u] = Ys[k * LD + tr + 16 * u];
#pragma unroll
      for (int v = 0; v < 4; ++v) yy[v] = Ys[k * LD + tc + 16 * v];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (u < nvb) {
#pragma unroll
          for (int v = 0; v < 4; ++v)
            if (v <= u) acc[u][v] = fma(-yx[u], yy[v], acc[u][v]);   // lower triangle: blocks v <= u only
        }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        const int x = tr + 16 * u, y = tc + 16 * v;
        if (x < b && y <= x) Cout[(int64_t)x * b + y] = acc[u][v];
      }
  }
}

// The pull alone, on global memory (every pull level unless the fused kernel runs it): one thread per DESTINATION entry
// (ra, rc) of the parent's front index space [0, s + b)^2; it looks the entry up in every child through the inverse
// relative indices and adds the children's contributions in plan order.  Any number of children in ONE launch, no two
// threads share a destination: no atomics, bitwise reproducible.  Entries with ra in the boundary and rc in the own
// block are the mirror of a coupling entry and are not stored.
template <class T>
__global__ void __launch_bounds__(256)
pull_children_kernel(FrontFusedArgs a, T* __restrict__ A, T* __restrict__ Bt, const T* __restrict__ Call,
                     T* __restrict__ Cout) {
  pdl_trigger();
  pdl_wait();
  const int f = blockIdx.y, s = a.s, b = a.b, span = s + b;
  const int g = a.f0 + f;
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= (int64_t)span * span) return;
  const int ra = (int)(q / span), rc = (int)(q - (int64_t)ra * span);
  if (ra >= s && rc < s) return;
  T acc = Sc<T>::zero();
  bool hit = false;
  for (int qc = a.child_ptr[g]; qc < a.child_ptr[g + 1]; ++qc) {
    const int ch = a.child_ids[qc];
    const int32_t* __restrict__ inv = a.inv_rel + a.f_invOff[ch];
    const int ia = inv[ra], ic = inv[rc];
    if (ia < 0 || ic < 0) continue;
    const int cb = a.f_b[ch];
    const T* __restrict__ U = Call + a.f_offC[ch];
    acc = add(acc, (ia >= ic) ? U[(int64_t)ia * cb + ic] : U[(int64_t)ic * cb + ia]);
    hit = true;
  }
  if (!hit) return;
  T* dst;
  if (ra < s) dst = (rc < s) ? A + ((int64_t)f * s * s + (int64_t)ra * s + rc)
                             : Bt + ((int64_t)f * s * b + (int64_t)ra * b + (rc - s));
  else dst = Cout + ((int64_t)f * b * b + (int64_t)(ra - s) * b + (rc - s));
  *dst = add(*dst, acc);
}

template <class T>
inline cudaError_t potrf_tile_prepare() {
  static bool done = false;
  if (done) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(potrf_trtri_tile_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)(3 * TILE * (TILE + 1) * sizeof(T)));
  if (e == cudaSuccess) done = true;
  return e;
}

// A helper stream with its own event ring: work that is independent of the caller's next kernels is pushed to
// `s` (order(from, s)) and joined back later (order(s, to)).  Valid under stream capture (fork / join branches).
struct SideLane {
  cudaStream_t s = nullptr;
  cudaEvent_t ev[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  int next = 0;
  cudaError_t init() {
    if (s) return cudaSuccess;
    cudaError_t e = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    for (int k = 0; k < 8 && e == cudaSuccess; ++k) e = cudaEventCreateWithFlags(&ev[k], cudaEventDisableTiming);
    return e;
  }
  void destroy() {
    if (!s) return;
    cudaStreamSynchronize(s);
    cudaStreamDestroy(s);
    for (auto e : ev)
      if (e) cudaEventDestroy(e);
    s = nullptr;
  }
  // `to` waits for everything enqueued on `from` so far
  cudaError_t order(cudaStream_t from, cudaStream_t to) {
    if (from == to) return cudaSuccess;
    cudaEvent_t e = ev[next];
    next = (next + 1) & 7;
    cudaError_t rc = cudaEventRecord(e, from);
    return rc == cudaSuccess ? cudaStreamWaitEvent(to, e, 0) : rc;
  }
};

// Blocked factorisation of `batch` n x n SPD / complex-symmetric matrices:
//   A  (destroyed: trailing updates are applied in place), L and Linv must be ZERO-initialised by the
//   caller above their block diagonal (they are treated as dense by the GEMMs that follow);
//   tmp: workspace of at least TILE * n elements per batch entry (stride sT).
//   lane (optional): the off-diagonal row blocks of Linv are not needed by the factorisation itself; with a
//   lane, row block t is computed on the lane's stream right after tile t, beside the panel / trailing update
//   of step t and the tiles that follow, and only the last row block remains after the last tile.
template <class T>
inline void potrf_trtri_blocked(cudaStream_t st, int64_t n, T* A, int64_t lda, int64_t sA, T* L,
                                int64_t ldl, int64_t sL, T* Linv, int64_t ldi, int64_t sI, T* tmp,
                                int64_t sT, int batch, int* info, SideLane* lane = nullptr) {
  if (n <= 0 || batch <= 0) return;
  potrf_tile_prepare<T>();
  const size_t smem = 3 * TILE * (TILE + 1) * sizeof(T);
  const T one = Sc<T>::one(), zero = Sc<T>::zero(), mone = neg(Sc<T>::one());
  const int64_t nt = (n + TILE - 1) / TILE;
  if (nt < 3 || !lane || !lane->s) lane = nullptr;
  cudaStream_t sl = lane ? lane->s : st;
  // Linv row-block s, columns [0, o): Linv[s,0:s] = -Linv_ss * (L[s,0:s] * Linv[0:s,0:s])
  auto linv_row = [&](int64_t s) {
    const int64_t o = s * TILE, bs = (n - o < TILE) ? n - o : TILE;
    gemm<T>(sl, 0, 0, bs, o, o, one, L + o * ldl, ldl, sL, Linv, ldi, sI, zero, tmp, o, sT, batch);
    gemm<T>(sl, 0, 0, bs, o, bs, mone, Linv + o * ldi + o, ldi, sI, tmp, o, sT, zero, Linv + o * ldi, ldi, sI,
            batch);
  };
  for (int64_t t = 0; t < nt; ++t) {
    const int64_t o = t * TILE, bt = (n - o < TILE) ? n - o : TILE;
    launch_pdl(potrf_trtri_tile_kernel<T>, dim3((unsigned)batch), dim3(256), smem, st, (int)bt,
               (const T*)(A + o * lda + o), lda, sA, L + o * ldl + o, ldl, sL, Linv + o * ldi + o, ldi, sI, info);
    CEIT_COUNT_LAUNCH();
    g_flops += (Sc<T>::is_complex ? 4.0 : 1.0) * (2.0 / 3.0) * (double)bt * bt * bt * batch;
    if (lane && t >= 1) {
      lane->order(st, sl);          // tile t (Linv_tt) and the panels L[t, 0:t] of the earlier steps
      linv_row(t);
    }
    const int64_t mr = n - o - bt;
    if (mr > 0) {
      const int64_t o2 = o + bt;
      // panel: L[o2:, o:o+bt] = A[o2:, o:o+bt] * Linv_tt^T
      gemm<T>(st, 0, 1, mr, bt, bt, one, A + o2 * lda + o, lda, sA, Linv + o * ldi + o, ldi, sI, zero,
              L + o2 * ldl + o, ldl, sL, batch);
      // trailing: A[o2:, o2:] -= Lp Lp^T
      gemm<T>(st, 0, 1, mr, mr, bt, mone, L + o2 * ldl + o, ldl, sL, L + o2 * ldl + o, ldl, sL, one,
              A + o2 * lda + o2, lda, sA, batch, /*lower_only=*/1);
    }
  }
  if (lane) {
    lane->order(sl, st);
  } else {
    for (int64_t s = 1; s < nt; ++s) linv_row(s);
  }
}

}  // namespace ceit
