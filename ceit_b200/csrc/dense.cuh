// Dense symmetric factorisation primitives: Cholesky L L^T = A (plain transpose; with the complex
// square root this also covers the complex-SYMMETRIC matrices of the non-zero-base path) together
// with the explicit triangular inverse Linv = L^-1, so that every triangular solve of the path
// becomes a tensor-core GEMM.  Batched over blockIdx.(x) with strides.
#pragma once
#include "gemm.cuh"

namespace ceit {

constexpr int TILE = 64;

// One CTA factors one n x n (n <= 64) tile held in shared memory.
//   A (read), L (write, lower incl. zeros above the diagonal), Linv (write, same shape).
// info: set to 1 if a pivot is not positive (real) / zero (complex).
template <class T>
__global__ void __launch_bounds__(256)
potrf_trtri_tile_kernel(int n, const T* __restrict__ A, int64_t lda, int64_t sA, T* __restrict__ L,
                        int64_t ldl, int64_t sL, T* __restrict__ Linv, int64_t ldi, int64_t sI,
                        int* __restrict__ info) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* S = reinterpret_cast<T*>(smem_raw);        // [64][65]
  T* X = S + TILE * (TILE + 1);                 // [64][65]
  __shared__ T rinv_s;
  const int tid = threadIdx.x;
  A += (int64_t)blockIdx.x * sA;
  L += (int64_t)blockIdx.x * sL;
  Linv += (int64_t)blockIdx.x * sI;
  for (int q = tid; q < n * n; q += 256) {
    int r = q / n, c = q - r * n;
    S[r * (TILE + 1) + c] = A[(int64_t)r * lda + c];
  }
  __syncthreads();
  // right-looking Cholesky on the lower triangle
  for (int j = 0; j < n; ++j) {
    if (tid == 0) {
      T d = S[j * (TILE + 1) + j];
      if (bad_pivot(d)) atomicExch(info, 1);
      d = sqrt_(d);
      S[j * (TILE + 1) + j] = d;
      rinv_s = inv_(d);
    }
    __syncthreads();
    T rinv = rinv_s;
    for (int i = j + 1 + tid; i < n; i += 256) S[i * (TILE + 1) + j] = mul(S[i * (TILE + 1) + j], rinv);
    __syncthreads();
    const int rem = n - j - 1;
    for (int q = tid; q < rem * rem; q += 256) {
      int a = q / rem, b = q - a * rem;
      if (b <= a) {
        int i = j + 1 + a, k = j + 1 + b;
        S[i * (TILE + 1) + k] = sub(S[i * (TILE + 1) + k], mul(S[i * (TILE + 1) + j], S[k * (TILE + 1) + j]));
      }
    }
    __syncthreads();
  }
  // Linv: column c by forward substitution, one thread per column
  if (tid < n) {
    const int c = tid;
    for (int r = 0; r < c; ++r) X[r * (TILE + 1) + c] = Sc<T>::zero();
    for (int r = c; r < n; ++r) {
      T s0 = (r == c) ? Sc<T>::one() : Sc<T>::zero();
      T s1 = Sc<T>::zero();
      int k = c;
      for (; k + 1 < r; k += 2) {
        s0 = sub(s0, mul(S[r * (TILE + 1) + k], X[k * (TILE + 1) + c]));
        s1 = sub(s1, mul(S[r * (TILE + 1) + k + 1], X[(k + 1) * (TILE + 1) + c]));
      }
      if (k < r) s0 = sub(s0, mul(S[r * (TILE + 1) + k], X[k * (TILE + 1) + c]));
      X[r * (TILE + 1) + c] = mul(add(s0, s1), inv_(S[r * (TILE + 1) + r]));
    }
  }
  __syncthreads();
  for (int q = tid; q < n * n; q += 256) {
    int r = q / n, c = q - r * n;
    T z = Sc<T>::zero();
    L[(int64_t)r * ldl + c] = (c <= r) ? S[r * (TILE + 1) + c] : z;
    Linv[(int64_t)r * ldi + c] = (c <= r) ? X[r * (TILE + 1) + c] : z;
  }
}

template <class T>
inline cudaError_t potrf_tile_prepare() {
  static bool done = false;
  if (done) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(potrf_trtri_tile_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)(2 * TILE * (TILE + 1) * sizeof(T)));
  if (e == cudaSuccess) done = true;
  return e;
}

// Blocked factorisation of `batch` n x n SPD / complex-symmetric matrices:
//   A  (destroyed: trailing updates are applied in place), L and Linv must be ZERO-initialised by the
//   caller above their block diagonal (they are treated as dense by the GEMMs that follow);
//   tmp: workspace of at least TILE * n elements per batch entry (stride sT).
template <class T>
inline void potrf_trtri_blocked(cudaStream_t st, int64_t n, T* A, int64_t lda, int64_t sA, T* L,
                                int64_t ldl, int64_t sL, T* Linv, int64_t ldi, int64_t sI, T* tmp,
                                int64_t sT, int batch, int* info) {
  if (n <= 0 || batch <= 0) return;
  potrf_tile_prepare<T>();
  const size_t smem = 2 * TILE * (TILE + 1) * sizeof(T);
  const T one = Sc<T>::one(), zero = Sc<T>::zero(), mone = neg(Sc<T>::one());
  const int64_t nt = (n + TILE - 1) / TILE;
  for (int64_t t = 0; t < nt; ++t) {
    const int64_t o = t * TILE, bt = (n - o < TILE) ? n - o : TILE;
    potrf_trtri_tile_kernel<T><<<batch, 256, smem, st>>>((int)bt, A + o * lda + o, lda, sA, L + o * ldl + o, ldl, sL,
                                                         Linv + o * ldi + o, ldi, sI, info);
    CEIT_COUNT_LAUNCH();
    const int64_t mr = n - o - bt;
    if (mr > 0) {
      const int64_t o2 = o + bt;
      // panel: L[o2:, o:o+bt] = A[o2:, o:o+bt] * Linv_tt^T
      gemm<T>(st, 0, 1, mr, bt, bt, one, A + o2 * lda + o, lda, sA, Linv + o * ldi + o, ldi, sI, zero,
              L + o2 * ldl + o, ldl, sL, batch);
      // trailing: A[o2:, o2:] -= Lp Lp^T
      gemm<T>(st, 0, 1, mr, mr, bt, mone, L + o2 * ldl + o, ldl, sL, L + o2 * ldl + o, ldl, sL, one,
              A + o2 * lda + o2, lda, sA, batch);
    }
  }
  // Linv row-block s, columns [0, o): Linv[s,0:s] = -Linv_ss * (L[s,0:s] * Linv[0:s,0:s])
  for (int64_t s = 1; s < nt; ++s) {
    const int64_t o = s * TILE, bs = (n - o < TILE) ? n - o : TILE;
    gemm<T>(st, 0, 0, bs, o, o, one, L + o * ldl, ldl, sL, Linv, ldi, sI, zero, tmp, o, sT, batch);
    gemm<T>(st, 0, 0, bs, o, bs, mone, Linv + o * ldi + o, ldi, sI, tmp, o, sT, zero, Linv + o * ldi, ldi, sI,
            batch);
  }
}

}  // namespace ceit
