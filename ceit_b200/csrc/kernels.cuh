// Path-specific kernels: element assembly (K1), dense scatter, Schur/Woodbury helpers and the
// per-(excitation, element) Woodbury kernel (K2).  All HBM-bound byte/f64 work: coalesced row reads,
// warp-level reductions, no atomics on the data path.
#pragma once
#include "gemm.cuh"

namespace ceit {

// ------------------------------------------------------------------------------------------------
// K1  element matrices -> CSR values, deterministic gather form.
// One thread per CSR slot sums the (element, a, b) contributions listed for it in ascending element
// order, recomputing the linear-triangle gradients from the node coordinates:
//   k_ab = perm_e (b_a b_b + c_a c_b) 4A  +  j omega var_e A (1/6 | 1/12)       (CEIT/EFEM.py:59-65)
//   K[n_a, n_b] += 2 k_ab                                                        (EFEM.py:61-62,66-67)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tri_geometry(const double* __restrict__ nodes, const int32_t* __restrict__ elem,
                                             int e, double& area, double b[3], double c[3], double& xb,
                                             double& yb) {
  const int n0 = elem[3 * e], n1 = elem[3 * e + 1], n2 = elem[3 * e + 2];
  const double2 p0 = reinterpret_cast<const double2*>(nodes)[n0];
  const double2 p1 = reinterpret_cast<const double2*>(nodes)[n1];
  const double2 p2 = reinterpret_cast<const double2*>(nodes)[n2];
  const double det = p0.x * (p1.y - p2.y) - p0.y * (p1.x - p2.x) + (p1.x * p2.y - p2.x * p1.y);
  const double id = 1.0 / det;
  area = fabs(det * 0.5);
  b[0] = (p1.y - p2.y) * id;
  b[1] = (p2.y - p0.y) * id;
  b[2] = (p0.y - p1.y) * id;
  c[0] = (p2.x - p1.x) * id;
  c[1] = (p0.x - p2.x) * id;
  c[2] = (p1.x - p0.x) * id;
  xb = (p0.x + p1.x + p2.x) / 3.0;
  yb = (p0.y + p1.y + p2.y) / 3.0;
}

__global__ void __launch_bounds__(256)
assemble_csr_kernel(int nnz, const int32_t* __restrict__ src_ptr, const int32_t* __restrict__ src,
                    const double* __restrict__ nodes, const int32_t* __restrict__ elem,
                    const double* __restrict__ perm, const double* __restrict__ var, double omega,
                    cd* __restrict__ vals) {
  pdl_trigger();
  pdl_wait();
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nnz) return;
  double re = 0.0, im = 0.0;
  for (int q = src_ptr[s]; q < src_ptr[s + 1]; ++q) {
    const int code = src[q];
    const int e = code / 9, ab = code - 9 * e, a = ab / 3, bb = ab - 3 * a;
    double area, b[3], c[3], xb, yb;
    tri_geometry(nodes, elem, e, area, b, c, xb, yb);
    const double kst = perm[e] * (b[a] * b[bb] + c[a] * c[bb]) * (4.0 * area);
    const double kms = (a == bb) ? (omega * var[e] * area / 6.0) : (omega * var[e] * area / 12.0);
    re += 2.0 * kst;   // added to [i][j] and [j][i]; the pattern lists both orders
    im += 2.0 * kms;
  }
  vals[s] = mk(re, im);
}

// K1, element-parallel form (default).  Pass 1: one thread per ELEMENT reads its connectivity (12 contiguous bytes,
// coalesced across the warp), gathers the three node coordinates as 16-byte loads, derives the triangle geometry ONCE
// and stores the nine (a, b) contributions as 16-byte (re, im) pairs at their position in the slot-sorted contribution
// list (`elem_dst`, the inverse of the gather list `src`): the contributions of a CSR slot end up contiguous, in
// ascending element order.  Pass 2: one thread per CSR slot sums its contiguous run (a segmented reduction over a
// coalesced stream, the same order as the gather form, so the result does not depend on the launch geometry).
// Against the gather form this removes the ~9x redundant geometry and 6/7 of the scattered 16-byte gathers.
__global__ void __launch_bounds__(256)
assemble_elem_kernel(int E, const double* __restrict__ nodes, const int32_t* __restrict__ elem,
                     const double* __restrict__ perm, const double* __restrict__ var, double omega,
                     const int32_t* __restrict__ elem_dst, double2* __restrict__ contrib) {
  pdl_trigger();
  pdl_wait();
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  double area, b[3], c[3], xb, yb;
  tri_geometry(nodes, elem, e, area, b, c, xb, yb);
  const double pe = perm[e], a4 = 4.0 * area;
  const double ms = omega * var[e] * area;
  const double md = 2.0 * (ms / 6.0), mo = 2.0 * (ms / 12.0);
  const int32_t* d = elem_dst + 9 * (int64_t)e;
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int bb = a; bb < 3; ++bb) {
      const double kst = 2.0 * (pe * fma(b[a], b[bb], c[a] * c[bb]) * a4);   // symmetric in (a, b) bit for bit
      const double2 v = make_double2(kst, a == bb ? md : mo);
      contrib[d[3 * a + bb]] = v;
      if (a != bb) contrib[d[3 * bb + a]] = v;
    }
}

__global__ void __launch_bounds__(256)
assemble_reduce_kernel(int nnz, const int32_t* __restrict__ src_ptr, const double2* __restrict__ contrib,
                       cd* __restrict__ vals) {
  pdl_trigger();
  pdl_wait();
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nnz) return;
  double re = 0.0, im = 0.0;
  const int q1 = src_ptr[s + 1];
  for (int q = src_ptr[s]; q < q1; ++q) {
    const double2 v = contrib[q];
    re += v.x;
    im += v.y;
  }
  vals[s] = mk(re, im);
}

__global__ void __launch_bounds__(256)
elem_param_kernel(int E, const double* __restrict__ nodes, const int32_t* __restrict__ elem,
                  double* __restrict__ param, double* __restrict__ area_out) {
  pdl_trigger();
  pdl_wait();
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  double area, b[3], c[3], xb, yb;
  tri_geometry(nodes, elem, e, area, b, c, xb, yb);
  if (area_out) area_out[e] = area;
  if (param) {
    double* p = param + 9 * (int64_t)e;
    p[0] = area;
    p[1] = b[0]; p[2] = b[1]; p[3] = b[2];
    p[4] = c[0]; p[5] = c[1]; p[6] = c[2];
    p[7] = xb; p[8] = yb;
  }
}

// Mesh model on the device (MeshObj.initialize_parameters / calc_electrode_elements / calc_detection_elements,
// CEIT/models/mesh.py:41-75,90-113,127-155): per element [A, b0..2, c0..2, xbar, ybar], membership of the closed square
// of every electrode (centroid test, mesh.py:104-107) and of the open detection square (mesh.py:148-151).  The
// comparisons use exactly the reference's expressions (centre +- radius against the centroid), so the sets are
// bit-identical to the host model's; int64 connectivity as MeshObj holds it.
__global__ void __launch_bounds__(256)
mesh_model_kernel(int64_t E, int L, const double* __restrict__ nodes, const int64_t* __restrict__ elem,
                  const double* __restrict__ centers, double radius, double det_bound, double* __restrict__ param,
                  uint8_t* __restrict__ in_electrode, uint8_t* __restrict__ in_detection) {
  pdl_trigger();
  pdl_wait();
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const double2* nd = reinterpret_cast<const double2*>(nodes);
  const double2 p0 = nd[elem[3 * e]], p1 = nd[elem[3 * e + 1]], p2 = nd[elem[3 * e + 2]];
  // explicitly rounded products and sums (no FMA contraction): the same roundings as the host expressions of
  // MeshObj.initialize_parameters, so elem_param agrees bit for bit (det cancels ~1e5 : 1 on a 0.8 m sheet)
  const double det = __dadd_rn(__dsub_rn(__dmul_rn(p0.x, __dsub_rn(p1.y, p2.y)), __dmul_rn(p0.y, __dsub_rn(p1.x, p2.x))),
                               __dsub_rn(__dmul_rn(p1.x, p2.y), __dmul_rn(p2.x, p1.y)));
  const double xb = __dadd_rn(__dadd_rn(p0.x, p1.x), p2.x) / 3.0, yb = __dadd_rn(__dadd_rn(p0.y, p1.y), p2.y) / 3.0;
  if (param) {
    double* p = param + 9 * e;
    p[0] = fabs(det / 2);
    p[1] = (p1.y - p2.y) / det; p[2] = (p2.y - p0.y) / det; p[3] = (p0.y - p1.y) / det;
    p[4] = (p2.x - p1.x) / det; p[5] = (p0.x - p2.x) / det; p[6] = (p1.x - p0.x) / det;
    p[7] = xb; p[8] = yb;
  }
  if (in_detection) in_detection[e] = (fabs(xb) < det_bound && fabs(yb) < det_bound) ? 1 : 0;
  if (in_electrode)
    for (int m = 0; m < L; ++m) {
      const double cx = centers[2 * m], cy = centers[2 * m + 1];
      in_electrode[(int64_t)m * E + e] =
          (__dadd_rn(cx, radius) >= xb && xb >= __dsub_rn(cx, radius) && __dadd_rn(cy, radius) >= yb &&
           yb >= __dsub_rn(cy, radius)) ? 1 : 0;
    }
}

// CSR values -> dense block storage (diag blocks, sub-diagonal blocks, K0U, KUU), type T.
template <class T>
__global__ void __launch_bounds__(256)
scatter_dense_kernel(int nnz, const cd* __restrict__ vals, const int64_t* __restrict__ dst,
                     const uint8_t* __restrict__ kind, T* __restrict__ Ad, T* __restrict__ Bs,
                     T* __restrict__ K0U, T* __restrict__ KUU, T* __restrict__ Aint, T* __restrict__ Bt) {
  pdl_trigger();
  pdl_wait();
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nnz) return;
  const int k = kind[s];
  if (k == 0) return;
  const T v = Sc<T>::from_cd(vals[s]);
  const int64_t d = dst[s];
  if (k == 1) Ad[d] = v;
  else if (k == 2) Bs[d] = v;
  else if (k == 3) K0U[d] = v;
  else if (k == 4) KUU[d] = v;
  else if (k == 5) Aint[d] = v;
  else Bt[d] = v;
}

// Domain-decomposed base stage: the update matrices (b x b, boundary-padded) of the sub-tree roots travel through export
// slots of `slot_sz` elements.  to_xg = 1: C storage -> slots [slot0, slot0 + gridDim.y) (this rank's roots);
// to_xg = 0: slots -> C storage for every slot outside [skip0, skip1) (the other ranks' roots, after the all-gather).
template <class T>
__global__ void __launch_bounds__(256)
export_copy_kernel(int slot0, int skip0, int skip1, const int32_t* __restrict__ exp_b, const int64_t* __restrict__ exp_offC,
                   int64_t slot_sz, T* __restrict__ C, T* __restrict__ xg, int to_xg) {
  pdl_trigger();
  pdl_wait();
  const int slot = slot0 + blockIdx.y;
  if (slot >= skip0 && slot < skip1) return;
  const int64_t b = exp_b[slot], n = b * b;
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  if (to_xg) xg[slot * slot_sz + q] = C[exp_offC[slot] + q];
  else C[exp_offC[slot] + q] = xg[slot * slot_sz + q];
}
template <class T>
__global__ void add_lower_kernel(int64_t n, const T* __restrict__ P, T* __restrict__ S) {
  pdl_trigger();
  pdl_wait();
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q < n) S[q] = add(S[q], P[q]);
}

// nested dissection helpers ------------------------------------------------------------------------
// identity on the diagonal of the unused (padding) slots of every interior block
template <class T>
__global__ void pad_identity_kernel(int P, int nI, const int32_t* __restrict__ node_at, T* __restrict__ Aint) {
  pdl_trigger();
  pdl_wait();
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= P * nI) return;
  if (node_at[q] < 0) {
    const int p = q / nI, a = q - p * nI;
    Aint[(int64_t)p * nI * nI + (int64_t)a * nI + a] = Sc<T>::one();
  }
}
// level-2 block storage -= sum of the interiors' Schur contributions (fixed order: deterministic)
template <class T>
__global__ void m2_gather_kernel(int64_t nd, const int64_t* __restrict__ dst, const int32_t* __restrict__ src_ptr,
                                 const int32_t* __restrict__ src, const T* __restrict__ Tm, int64_t szD,
                                 T* __restrict__ Ad, T* __restrict__ Bs) {
  pdl_trigger();
  pdl_wait();
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nd) return;
  T acc = Sc<T>::zero();
  for (int s = src_ptr[q]; s < src_ptr[q + 1]; ++s) acc = add(acc, Tm[src[s]]);
  const int64_t d = dst[q];
  if (d < szD) Ad[d] = sub(Ad[d], acc);
  else Bs[d - szD] = sub(Bs[d - szD], acc);
}
// r_C[c,:] = K0U[c,:] - sum over boundary occurrences of row (p,a) of G1; one warp per separator node
template <class T>
__global__ void __launch_bounds__(256)
rc_gather_kernel(int64_t nC, int nU, const int32_t* __restrict__ rc_ptr, const int32_t* __restrict__ rc_src,
                 const T* __restrict__ G1, const T* __restrict__ K0U_C, T* __restrict__ Z_C) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t c = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (c >= nC) return;
  for (int k = lane; k < nU; k += 32) {
    T acc = K0U_C[c * nU + k];
    for (int s = rc_ptr[c]; s < rc_ptr[c + 1]; ++s) acc = sub(acc, G1[(int64_t)rc_src[s] * nU + k]);
    Z_C[c * nU + k] = acc;
  }
}
// G1[(p,a),:] = X_C[bnd_loc[p,a],:] (zero rows for padding)
template <class T>
__global__ void __launch_bounds__(256)
bnd_rows_gather_kernel(int64_t nrows, int nU, const int32_t* __restrict__ bnd_loc, const T* __restrict__ X_C,
                       T* __restrict__ G1) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= nrows) return;
  const int32_t c = bnd_loc[r];
  for (int k = lane; k < nU; k += 32) G1[r * nU + k] = (c < 0) ? Sc<T>::zero() : X_C[(int64_t)c * nU + k];
}

// ---- multi-level nested dissection (factor_tree.h) ------------------------------------------------------------------
// identity on the diagonal of every padding slot of the own blocks (host list of A offsets)
template <class T>
__global__ void pad_diag_list_kernel(int64_t n, const int64_t* __restrict__ dst, T* __restrict__ A) {
  pdl_trigger();
  pdl_wait();
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q < n) A[dst[q]] = Sc<T>::one();
}
// Extend-add of one level: the update matrix U_f (b x b, lower triangle valid) of every front whose sibling index is
// `pass` is added into its parent's own block A, coupling block Bt (own rows x boundary columns) and boundary block C
// through the relative indices rel (index of the boundary node in the parent's front index space: [0, ps) own slots,
// then boundary slots).  Siblings are separated by `pass` launches and rel is injective, so no two threads of a
// launch touch the same entry: no atomics, bitwise reproducible.  blockIdx.y = front of the level.
template <class T>
__global__ void __launch_bounds__(256)
extend_add_kernel(int b, int pass, const uint8_t* __restrict__ push, const int32_t* __restrict__ sib,
                  const int64_t* __restrict__ pA,
                  const int64_t* __restrict__ pB, const int64_t* __restrict__ pC, const int32_t* __restrict__ ps,
                  const int32_t* __restrict__ pb, const int32_t* __restrict__ rel, const T* __restrict__ U,
                  T* __restrict__ A, T* __restrict__ Bt, T* __restrict__ C) {
  pdl_trigger();
  pdl_wait();
  const int f = blockIdx.y;
  if (sib[f] != pass || pA[f] < 0 || !push[f]) return;   // fronts whose parent level pulls do not push
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= b * b) return;
  const int a = q / b, c = q - a * b;
  const int ra = rel[(int64_t)f * b + a], rc = rel[(int64_t)f * b + c];
  if (ra < 0 || rc < 0) return;
  const T* u = U + (int64_t)f * b * b;
  const T v = (a >= c) ? u[(int64_t)a * b + c] : u[(int64_t)c * b + a];
  const int s_ = ps[f], b_ = pb[f];
  if (ra < s_) {
    if (rc < s_) {
      T* d = A + pA[f] + (int64_t)ra * s_ + rc;
      *d = add(*d, v);
    } else {
      T* d = Bt + pB[f] + (int64_t)ra * b_ + (rc - s_);
      *d = add(*d, v);
    }
  } else if (rc >= s_) {
    T* d = C + pC[f] + (int64_t)(ra - s_) * b_ + (rc - s_);
    *d = add(*d, v);
  }
}
// The transpose of extend-add for the selected inverse: Sigma_bb(f)[a][c] = entry (rel a, rel c) of the parent's
// selected inverse [[Sigma_own, V], [V^T, Sigma_bb(parent)]]; zero on padding / for roots.  Out = C block of f.
template <class T>
__global__ void __launch_bounds__(256)
extend_gather_kernel(int b, const int64_t* __restrict__ pA, const int64_t* __restrict__ pB,
                     const int64_t* __restrict__ pC, const int32_t* __restrict__ ps, const int32_t* __restrict__ pb,
                     const int32_t* __restrict__ rel, const T* __restrict__ Sg, const T* __restrict__ V,
                     const T* __restrict__ Call, T* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const int f = blockIdx.y;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= b * b) return;
  const int a = q / b, c = q - a * b;
  const int ra = rel[(int64_t)f * b + a], rc = rel[(int64_t)f * b + c];
  T v = Sc<T>::zero();
  if (ra >= 0 && rc >= 0 && pA[f] >= 0) {
    const int s_ = ps[f], b_ = pb[f];
    if (ra < s_ && rc < s_) v = Sg[pA[f] + (int64_t)ra * s_ + rc];
    else if (ra < s_) v = V[pB[f] + (int64_t)ra * b_ + (rc - s_)];
    else if (rc < s_) v = V[pB[f] + (int64_t)rc * b_ + (ra - s_)];
    else v = Call[pC[f] + (int64_t)(ra - s_) * b_ + (rc - s_)];
  }
  out[(int64_t)f * b * b + q] = v;
}
// R[dst[q], :] -= sum_s G[src[s], :] (destination-major lists: deterministic); blockIdx.x = destination row,
// blockIdx.y = chunk of 128 columns, one column per thread (all loads of a thread are independent)
template <class T>
__global__ void __launch_bounds__(128)
rhs_scatter_kernel(int64_t nd, int nU, const int32_t* __restrict__ dst, const int32_t* __restrict__ ptr,
                   const int32_t* __restrict__ src, const T* __restrict__ G, T* __restrict__ R) {
  pdl_trigger();
  pdl_wait();
  const int64_t q = blockIdx.x;
  const int k = blockIdx.y * 128 + threadIdx.x;
  if (q >= nd || k >= nU) return;
  T* r = R + (int64_t)dst[q] * nU + k;
  const int s0 = ptr[q], s1 = ptr[q + 1];
  T acc = *r;
  for (int s = s0; s < s1; ++s) acc = sub(acc, G[(int64_t)src[s] * nU + k]);
  *r = acc;
}

// Xh[c,:] = XP[row_map[c],:]: drop the padding slots of the interiors (one warp per row)
template <class T>
__global__ void __launch_bounds__(256)
compact_rows_kernel(int64_t N0, int nU, const int32_t* __restrict__ row_map, const T* __restrict__ XP,
                    T* __restrict__ Xh) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= N0) return;
  const T* s = XP + (int64_t)row_map[r] * nU;
  for (int k = lane; k < nU; k += 32) Xh[r * nU + k] = s[k];
}

// out = -in
template <class T>
__global__ void negate_copy_kernel(int64_t n, const T* __restrict__ in, T* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q < n) out[q] = neg(in[q]);
}

// rows [N0, N) of Xh = -I
template <class T>
__global__ void fill_neg_identity_kernel(int nU, T* __restrict__ Xh_tail) {
  pdl_trigger();
  pdl_wait();
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nU * nU) return;
  const int r = q / nU, c = q - r * nU;
  Xh_tail[q] = (r == c) ? neg(Sc<T>::one()) : Sc<T>::zero();
}

// g0e[e][q] = selected-inverse entry for the element's node pair, 0 if either node is an electrode node
template <class T>
__global__ void gather_g0_kernel(int64_t n, const int64_t* __restrict__ off, const T* __restrict__ Sel,
                                 T* __restrict__ g0e) {
  pdl_trigger();
  pdl_wait();
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  const int64_t o = off[q];
  g0e[q] = (o < 0) ? Sc<T>::zero() : Sel[o];
}

// St[b] = S_U with the rows/cols of B_i replaced by identity (i = i0 + b)
template <class T>
__global__ void build_stilde_kernel(int nU, const T* __restrict__ SU, const uint8_t* __restrict__ inB, int i0,
                                    T* __restrict__ St) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.y;
  const uint8_t* mask = inB + (int64_t)(i0 + b) * nU;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nU * nU) return;
  const int r = q / nU, c = q - r * nU;
  T v = SU[q];
  if (mask[r] || mask[c]) v = (r == c) ? Sc<T>::one() : Sc<T>::zero();
  St[(int64_t)b * nU * nU + q] = v;
}

// zero rows/cols of B_i in Sinv[b]
template <class T>
__global__ void mask_sinv_kernel(int nU, const uint8_t* __restrict__ inB, int i0, T* __restrict__ Sinv) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.y;
  const uint8_t* mask = inB + (int64_t)(i0 + b) * nU;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nU * nU) return;
  const int r = q / nU, c = q - r * nU;
  if (mask[r] || mask[c]) Sinv[(int64_t)b * nU * nU + q] = Sc<T>::zero();
}

// ------------------------------------------------------------------------------------------------
// Per-excitation Schur stage as a rank-(b+1) update of ONE inverse:
//   M = S_U + gamma e e^T (e = 1/sqrt(nU): S_U 1 = 0 on the real base),  T = M^-1,  Y_T = Xh T,
//   Sinv_i = T - T[:,B] C_i + alpha q q^T,  C_i = T_BB^-1 T[B,:],  q = T e - T[:,B] C_i e (zero on B),
//   alpha = gamma / (1 - gamma e.q),   Yh_i = Y_T - Y_T[:,B] D_i + alpha (Y_T e) q^T,
//   D_i = C_i + alpha (C_i e) q^T.   B = B_i is the contiguous range [lo, lo+b) of U.
// ------------------------------------------------------------------------------------------------
template <class T>
__global__ void trace_gamma_kernel(int nU, const T* __restrict__ SU, double* __restrict__ gamma) {
  pdl_trigger();
  pdl_wait();
  __shared__ double red[256];
  double s = 0.0;
  for (int k = threadIdx.x; k < nU; k += 256) s += to_cd(SU[(int64_t)k * nU + k]).x;
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) gamma[0] = red[0] / nU;
}
template <class T>
__global__ void add_rank1_e_kernel(int nU, const T* __restrict__ SU, const double* __restrict__ gamma,
                                   T* __restrict__ M) {
  pdl_trigger();
  pdl_wait();
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nU * nU) return;
  M[q] = add(SU[q], Sc<T>::from_real(gamma[0] / nU));
}
// out[r] = (1/sqrt(nU)) sum_k A[r,k]; one warp per row
template <class T>
__global__ void __launch_bounds__(256)
rowsum_e_kernel(int64_t rows, int nU, const T* __restrict__ A, T* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows) return;
  T acc = Sc<T>::zero();
  for (int k = lane; k < nU; k += 32) acc = add(acc, A[r * nU + k]);
  cd a = to_cd(acc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a.x += __shfl_xor_sync(0xffffffffu, a.x, o);
    a.y += __shfl_xor_sync(0xffffffffu, a.y, o);
  }
  if (lane == 0) out[r] = Sc<T>::from_cd(scale(rsqrt((double)nU), a));
}
// TBB[b] = T[B_i, B_i] padded to bp x bp with identity
template <class T>
__global__ void gather_tbb_kernel(int nU, int bp, const T* __restrict__ Tm, const int32_t* __restrict__ ue_ptr,
                                  int i0, T* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const int i = i0 + blockIdx.y;
  const int lo = ue_ptr[i], b = ue_ptr[i + 1] - lo;
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= bp * bp) return;
  const int a = q / bp, c = q - a * bp;
  T v = (a == c) ? Sc<T>::one() : Sc<T>::zero();
  if (a < b && c < b) v = Tm[(int64_t)(lo + a) * nU + lo + c];
  out[(int64_t)blockIdx.y * bp * bp + q] = v;
}
// one CTA per excitation: Dcat[b] ((bp+1) x nU): rows j < b = D_i, row bp = alpha q, other rows 0.
// The b rows T[B_i,:] are staged in shared memory once (cp.async); a thread owns columns k = tid, tid + 256, ...
// and forms C[:,k] = T_BB^-1 T[B,k] in registers (8 rows at a time), so the kernel has no global-latency chains.
template <class T>
__global__ void __launch_bounds__(256)
excite_small_kernel(int nU, int bp, const T* __restrict__ Tm, const T* __restrict__ te,
                    const T* __restrict__ LinvB, const int32_t* __restrict__ ue_ptr, int i0,
                    const double* __restrict__ gamma, T* __restrict__ Dcat, int stage_rows) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* Li = reinterpret_cast<T*>(smem_raw);   // [bp][bp]
  T* Bi = Li + bp * bp;                     // [bp][bp]  T_BB^-1
  T* ce = Bi + bp * bp;                     // [bp]
  T* q = ce + bp;                           // [nU]
  T* red = q + nU;                          // [256]
  T* cep = red + 256;                       // [8][bp]   per-warp partial row sums of C
  T* Tsm = cep + 8 * bp;                    // [b][nU]   T[B_i, :] (only when stage_rows: it must fit)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int i = i0 + blockIdx.x;
  const int lo = ue_ptr[i], b = ue_ptr[i + 1] - lo;
  T* D = Dcat + (int64_t)blockIdx.x * (bp + 1) * nU;
  const T* Lg = LinvB + (int64_t)blockIdx.x * bp * bp;
  const T* Ts = stage_rows ? Tsm : Tm + (int64_t)lo * nU;                                            // rows lo..lo+b-1
  if (stage_rows)
    for (int t = tid; t < b * nU; t += 256) cp_async_elem(Tsm + t, Tm + (int64_t)lo * nU + t, true);
  cp_async_commit();
  for (int t = tid; t < bp * bp; t += 256) Li[t] = Lg[t];
  __syncthreads();
  for (int t = tid; t < bp * bp; t += 256) {
    const int a = t / bp, c = t - a * bp;
    T acc = Sc<T>::zero();
    for (int k = (a > c ? a : c); k < bp; ++k) acc = fma_(Li[k * bp + a], Li[k * bp + c], acc);
    Bi[t] = acc;
  }
  cp_async_wait<0>();
  __syncthreads();
  // C = T_BB^-1 T[B,:]  (stored in D for now); rows >= b are zero.  ce = C e from per-warp partial sums.
  for (int j0 = 0; j0 < bp; j0 += 8) {
    T rs[8];
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) rs[jj] = Sc<T>::zero();
    for (int k = tid; k < nU; k += 256) {
      T acc[8];
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) acc[jj] = Sc<T>::zero();
      if (j0 < b) {
        for (int m = 0; m < b; ++m) {
          const T tm = Ts[m * nU + k];
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) acc[jj] = fma_(Bi[(j0 + jj) * bp + m], tm, acc[jj]);   // rows >= b of Bi: identity pad
        }
      }
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const T v = (j0 + jj < b) ? acc[jj] : Sc<T>::zero();
        D[(int64_t)(j0 + jj) * nU + k] = v;
        rs[jj] = add(rs[jj], v);
      }
    }
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      cd a = to_cd(rs[jj]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        a.x += __shfl_xor_sync(0xffffffffu, a.x, o);
        a.y += __shfl_xor_sync(0xffffffffu, a.y, o);
      }
      if (lane == 0) cep[warp * bp + j0 + jj] = Sc<T>::from_cd(a);
    }
  }
  __syncthreads();
  const double se = rsqrt((double)nU);
  if (tid < bp) {
    T a = Sc<T>::zero();
#pragma unroll
    for (int w = 0; w < 8; ++w) a = add(a, cep[w * bp + tid]);
    ce[tid] = Sc<T>::from_cd(scale(se, to_cd(a)));
  }
  __syncthreads();
  // q = T e - T[:,B] ce (T symmetric: T[k, lo+j] = T[lo+j, k]); zero on B
  T part = Sc<T>::zero();
  for (int k = tid; k < nU; k += 256) {
    T v = te[k];
    for (int j = 0; j < b; ++j) v = sub(v, mul(Ts[j * nU + k], ce[j]));
    if (k >= lo && k < lo + b) v = Sc<T>::zero();
    q[k] = v;
    part = add(part, v);
  }
  red[tid] = part;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) red[tid] = add(red[tid], red[tid + o]);
    __syncthreads();
  }
  const T eq = Sc<T>::from_cd(scale(se, to_cd(red[0])));
  const double g = gamma[0];
  const T alpha = mul(Sc<T>::from_real(g), inv_(sub(Sc<T>::one(), scale(g, eq))));
  // D = C + alpha (C e) q^T: every thread revisits the (j, k) entries it wrote itself
  for (int k = tid; k < nU; k += 256) {
    const T aq = mul(alpha, q[k]);
    for (int j2 = 0; j2 < b; ++j2) D[(int64_t)j2 * nU + k] = add(D[(int64_t)j2 * nU + k], mul(ce[j2], aq));
    D[(int64_t)bp * nU + k] = aq;
  }
}
// out[b][n,k] = src[n,k] - sum_{j<b} src[n, lo+j] D[j,k] + yv[n] D[bp,k]; columns (and, if zero_rows, rows)
// of B_i are set to exactly zero.  CTA = one excitation x `passes` tiles of RT rows x one column chunk: D_i is
// staged in shared memory once per CTA.  The 8 warps form WC column groups x 8/WC row groups: a thread owns
// CPT columns (32 WC apart) of RH rows (8/WC apart), so a chunk is 32 WC CPT columns wide and the host picks
// (WC, CPT, number of chunks) with the least padding (nU = 544: 2 x 3 x 3 chunks = 576; nU = 264: 1 x 3 x 3 = 288).
template <class T, int CPT, int WC>
__global__ void __launch_bounds__(256)
rank_update_kernel(int64_t rows, int nU, int bp, const T* __restrict__ src, const T* __restrict__ yv,
                   const T* __restrict__ Dcat, const int32_t* __restrict__ ue_ptr, int i0, int zero_rows,
                   T* __restrict__ out, int64_t out_stride, int passes, int cw) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int RH = Sc<T>::is_complex ? 4 : 8;   // rows per thread
  constexpr int TXN = 32 * WC;                     // threads across the columns
  constexpr int NY = 256 / TXN;                    // row groups
  constexpr int RT = NY * RH;                      // rows per pass (power of two)
  constexpr int CWP = TXN * CPT;                   // padded chunk width: no column predicates in the inner loop
  const int tid = threadIdx.x, tx = tid % TXN, ty = tid / TXN;
  const int i = i0 + blockIdx.x;      // excitation fastest: the CTAs that share a row tile of src run together
  const int lo = ue_ptr[i], b = ue_ptr[i + 1] - lo;
  const int k_lo = blockIdx.z * cw, kw = min(cw, nU - k_lo);
  // compact copies: rows 0..b-1 of D_i and the alpha q row (stored as row b); As likewise [j][RT]
  T* Ds = reinterpret_cast<T*>(smem_raw);   // [b+1][CWP], zero beyond kw
  T* As = Ds + (int64_t)(bp + 1) * CWP;     // [b+1][RT]
  const T* D = Dcat + (int64_t)blockIdx.x * (bp + 1) * nU + k_lo;
  for (int j = 0; j <= b; ++j) {          // cp.async: all rows in flight at once, zero fill beyond kw
    const T* dr = D + (int64_t)(j < b ? j : bp) * nU;
    for (int k = tid; k < CWP; k += 256) cp_async_elem(Ds + j * CWP + k, dr + (k < kw ? k : 0), k < kw);
  }
  cp_async_commit();
  bool keep[CPT];                       // column inside the chunk and not a Dirichlet column of this excitation
#pragma unroll
  for (int c = 0; c < CPT; ++c) {
    const int k = tx + TXN * c, kg = k_lo + k;
    keep[c] = k < kw && !(kg >= lo && kg < lo + b);
  }
  T* o = out + (int64_t)blockIdx.x * out_stride + k_lo + tx;
  const T* sbase = src + k_lo + tx;
  for (int ps = 0; ps < passes; ++ps) {
    const int64_t n0 = ((int64_t)blockIdx.y * passes + ps) * RT;
    if (n0 >= rows) break;
    __syncthreads();                        // As of the previous pass consumed
    for (int t = tid; t < RT * (b + 1); t += 256) {
      const int j = t / RT, r = t % RT;     // RT is a power of two
      const int64_t n = n0 + r, nc = n < rows ? n : 0;
      cp_async_elem(As + t, (j < b) ? src + nc * nU + lo + j : yv + nc, n < rows);
    }
    cp_async_commit();
    T acc[RH][CPT];
    {
      const T* sp = sbase + (n0 + ty) * nU;
#pragma unroll
      for (int r = 0; r < RH; ++r) {
        const bool rok = n0 + NY * r + ty < rows;
#pragma unroll
        for (int c = 0; c < CPT; ++c) acc[r][c] = (rok && tx + TXN * c < kw) ? sp[TXN * c] : Sc<T>::zero();
        sp += NY * (int64_t)nU;
      }
    }
    cp_async_wait<0>();
    __syncthreads();
    {
      const T* dp = Ds + tx;
      const T* ap = As + ty;
#pragma unroll 2
      for (int j = 0; j < b; ++j) {
        T dv[CPT];
#pragma unroll
        for (int c = 0; c < CPT; ++c) dv[c] = dp[TXN * c];
#pragma unroll
        for (int r = 0; r < RH; ++r) {
          const T a = ap[NY * r];
#pragma unroll
          for (int c = 0; c < CPT; ++c) acc[r][c] = sub(acc[r][c], mul(a, dv[c]));
        }
        dp += CWP;
        ap += RT;
      }
      // + y_e (alpha q)^T
#pragma unroll
      for (int r = 0; r < RH; ++r) {
        const T a = ap[NY * r];
#pragma unroll
        for (int c = 0; c < CPT; ++c) acc[r][c] = fma_(a, dp[TXN * c], acc[r][c]);
      }
    }
    {
      T* op = o + (n0 + ty) * nU;
#pragma unroll
      for (int r = 0; r < RH; ++r) {
        const int64_t n = n0 + NY * r + ty;
        if (n < rows) {
          const bool rowB = zero_rows && (n >= lo && n < lo + b);
#pragma unroll
          for (int c = 0; c < CPT; ++c)
            if (tx + TXN * c < kw) op[TXN * c] = (keep[c] && !rowB) ? acc[r][c] : Sc<T>::zero();
        }
        op += NY * (int64_t)nU;
      }
    }
  }
}

// phiU[b] = -Sinv[b] * (sum_{u in B_i} SU[:,u]);  phiU[B_i] = 1, followed by one step of iterative
// refinement against S_U itself (stages 2, 3) so that phi0 does not inherit the rounding of Sinv.
// One warp per row.
template <class T>
__global__ void __launch_bounds__(256)
phi_u_kernel(int nU, const T* __restrict__ SU, const T* __restrict__ Sinv, const uint8_t* __restrict__ inB,
             int i0, T* __restrict__ tvec, T* __restrict__ phi0, int64_t N, int64_t N0, int stage) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.y;
  const uint8_t* mask = inB + (int64_t)(i0 + b) * nU;
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= nU) return;
  T acc = Sc<T>::zero();
  if (stage == 0) {  // tvec[row] = sum_{u in B_i} SU[row][u]
    const T* r = SU + (int64_t)row * nU;
    for (int u = lane; u < nU; u += 32)
      if (mask[u]) acc = add(acc, r[u]);
  } else if (stage == 1 || stage == 3) {   // sum_k Sinv[row][k] tvec[k]
    const T* r = Sinv + (int64_t)b * nU * nU + (int64_t)row * nU;
    const T* tv = tvec + (int64_t)b * nU;
    for (int k = lane; k < nU; k += 32) acc = fma_(r[k], tv[k], acc);
  } else {                                 // stage 2: residual of S_RR phi_R = -S_RB 1 with the current phi
    const T* r = SU + (int64_t)row * nU;
    const T* pu = phi0 + (int64_t)b * N + N0;
    for (int k = lane; k < nU; k += 32) acc = fma_(r[k], pu[k], acc);
  }
  cd a = to_cd(acc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a.x += __shfl_xor_sync(0xffffffffu, a.x, o);
    a.y += __shfl_xor_sync(0xffffffffu, a.y, o);
  }
  if (lane == 0) {
    T* pu = phi0 + (int64_t)b * N + N0;
    if (stage == 0) tvec[(int64_t)b * nU + row] = Sc<T>::from_cd(a);
    else if (stage == 1) pu[row] = mask[row] ? Sc<T>::one() : neg(Sc<T>::from_cd(a));
    else if (stage == 2) tvec[(int64_t)b * nU + row] = mask[row] ? Sc<T>::zero() : neg(Sc<T>::from_cd(a));
    else if (!mask[row]) pu[row] = add(pu[row], Sc<T>::from_cd(a));      // stage 3: one refinement step
  }
}

// phi0[b][row] = -sum_u X[row][u] phiU[b][u] for rows in F0. One warp per row.
template <class T>
__global__ void __launch_bounds__(256)
phi_f_kernel(int64_t N0, int nU, const T* __restrict__ X, T* __restrict__ phi0, int64_t N) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= N0) return;
  const T* r = X + row * nU;
  const T* pu = phi0 + (int64_t)b * N + N0;
  T acc = Sc<T>::zero();
  for (int u = lane; u < nU; u += 32) acc = fma_(r[u], pu[u], acc);
  cd a = to_cd(acc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a.x += __shfl_xor_sync(0xffffffffu, a.x, o);
    a.y += __shfl_xor_sync(0xffffffffu, a.y, o);
  }
  if (lane == 0) phi0[(int64_t)b * N + row] = neg(Sc<T>::from_cd(a));
}

// electrode amplitudes of the unperturbed solution: base rows (m != i) and/or all L electrode values
template <class T>
__global__ void baseline_kernel(int L, int nU, int64_t N, int64_t N0, const T* __restrict__ phi0,
                                const int32_t* __restrict__ we_ptr, const int32_t* __restrict__ we_u,
                                const double* __restrict__ we_w, int i0, double* __restrict__ base,
                                double* __restrict__ electrode_pot) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.x;
  const int i = i0 + b;
  const T* pu = phi0 + (int64_t)b * N + N0;
  for (int m = threadIdx.x; m < L; m += blockDim.x) {
    double s = 0.0;
    for (int q = we_ptr[m]; q < we_ptr[m + 1]; ++q) s = fma(we_w[q], abs_(pu[we_u[q]]), s);
    if (electrode_pot) electrode_pot[(int64_t)b * L + m] = s;
    if (base && m != i) base[(int64_t)i * (L - 1) + m - (m > i ? 1 : 0)] = fabs(s);
  }
}

// node / element amplitudes in ORIGINAL numbering (EFEM.sync_back_potential, FEMBasic.py:118-127)
template <class T>
__global__ void node_abs_kernel(int64_t N, const T* __restrict__ phi0, const int32_t* __restrict__ pos,
                                double* __restrict__ node_abs) {
  pdl_trigger();
  pdl_wait();
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v < N) node_abs[v] = abs_(phi0[pos[v]]);
}
__global__ void elem_mean_kernel(int64_t E, const int32_t* __restrict__ elem, const double* __restrict__ node_abs,
                                 double* __restrict__ elem_pot) {
  pdl_trigger();
  pdl_wait();
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e < E) elem_pot[e] = (node_abs[elem[3 * e]] + node_abs[elem[3 * e + 1]] + node_abs[elem[3 * e + 2]]) / 3.0;
}

// ------------------------------------------------------------------------------------------------
// K2  Woodbury kernel: one warp per (excitation, element).
//   G   = G0e + Yh[p] Xh[p]^T                         (3x3, = Kff^-1 on the element's nodes)
//   y   = (I + D G)^-1 D phi0[p],  D = j s M,  s = 2 omega c A_j,  M = (1 + I)/12
//   dU  = Yh[p]^T y  (perturbation at the nU electrode nodes),  amp = |phi0_U + dU|
//   J[i(L-1)+cnt(m), j] = | sum_u w[m,u] amp_u |,  m != i           (CEIT/EJAC.py:127-132)
// Rows of electrode / Dirichlet nodes need no special case: Xh/Yh are padded so that the same
// formula holds.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ cd cmul(cd a, cd b) { return mul(a, b); }
__device__ __forceinline__ cd csub(cd a, cd b) { return sub(a, b); }
__device__ __forceinline__ cd cadd(cd a, cd b) { return add(a, b); }

__device__ __forceinline__ void solve3(const cd A[3][3], const cd r[3], cd y[3]) {
  // adjugate / Cramer; A = I + small, no pivoting needed
  const cd c00 = csub(cmul(A[1][1], A[2][2]), cmul(A[1][2], A[2][1]));
  const cd c01 = csub(cmul(A[1][2], A[2][0]), cmul(A[1][0], A[2][2]));
  const cd c02 = csub(cmul(A[1][0], A[2][1]), cmul(A[1][1], A[2][0]));
  const cd det = cadd(cadd(cmul(A[0][0], c00), cmul(A[0][1], c01)), cmul(A[0][2], c02));
  const cd idet = inv_(det);
  const cd c10 = csub(cmul(A[0][2], A[2][1]), cmul(A[0][1], A[2][2]));
  const cd c11 = csub(cmul(A[0][0], A[2][2]), cmul(A[0][2], A[2][0]));
  const cd c12 = csub(cmul(A[0][1], A[2][0]), cmul(A[0][0], A[2][1]));
  const cd c20 = csub(cmul(A[0][1], A[1][2]), cmul(A[0][2], A[1][1]));
  const cd c21 = csub(cmul(A[0][2], A[1][0]), cmul(A[0][0], A[1][2]));
  const cd c22 = csub(cmul(A[0][0], A[1][1]), cmul(A[0][1], A[1][0]));
  // inverse = adj / det, adj[i][j] = cofactor[j][i]
  y[0] = cmul(idet, cadd(cadd(cmul(c00, r[0]), cmul(c10, r[1])), cmul(c20, r[2])));
  y[1] = cmul(idet, cadd(cadd(cmul(c01, r[0]), cmul(c11, r[1])), cmul(c21, r[2])));
  y[2] = cmul(idet, cadd(cadd(cmul(c02, r[0]), cmul(c12, r[1])), cmul(c22, r[2])));
}

template <class T>
__global__ void __launch_bounds__(256)
woodbury_kernel(int e_from, int e_to, int L, int nU, int64_t N, int64_t N0,
                const int32_t* __restrict__ elem_pos, const double* __restrict__ area,
                const T* __restrict__ g0e, const T* __restrict__ Xh, const T* __restrict__ Yh,
                const T* __restrict__ phi0, const int32_t* __restrict__ we_ptr,
                const int32_t* __restrict__ we_u, const double* __restrict__ we_w, int i0, double s_coef,
                double* __restrict__ J, int64_t ldJ) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warps = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* amp = reinterpret_cast<double*>(smem_raw) + (int64_t)warp * nU;
  const int b = blockIdx.y;
  const int i = i0 + b;
  const int j = e_from + blockIdx.x * warps + warp;
  if (j >= e_to) return;   // whole warp exits together; no block-level sync below
  const T* Y = Yh + (int64_t)b * N * nU;
  const T* ph = phi0 + (int64_t)b * N;
  const int p0 = elem_pos[3 * j], p1 = elem_pos[3 * j + 1], p2 = elem_pos[3 * j + 2];
  const T* y0 = Y + (int64_t)p0 * nU;
  const T* y1 = Y + (int64_t)p1 * nU;
  const T* y2 = Y + (int64_t)p2 * nU;
  const T* x0 = Xh + (int64_t)p0 * nU;
  const T* x1 = Xh + (int64_t)p1 * nU;
  const T* x2 = Xh + (int64_t)p2 * nU;
  // ---- six dot products: (0,0) (1,1) (2,2) (0,1) (1,2) (2,0)
  T d[6];
#pragma unroll
  for (int q = 0; q < 6; ++q) d[q] = Sc<T>::zero();
  for (int k = lane; k < nU; k += 32) {
    const T a0 = y0[k], a1 = y1[k], a2 = y2[k];
    const T b0 = x0[k], b1 = x1[k], b2 = x2[k];
    d[0] = fma_(a0, b0, d[0]);
    d[1] = fma_(a1, b1, d[1]);
    d[2] = fma_(a2, b2, d[2]);
    d[3] = fma_(a0, b1, d[3]);
    d[4] = fma_(a1, b2, d[4]);
    d[5] = fma_(a2, b0, d[5]);
  }
  cd G6[6];
#pragma unroll
  for (int q = 0; q < 6; ++q) {
    cd a = to_cd(d[q]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a.x += __shfl_xor_sync(0xffffffffu, a.x, o);
      a.y += __shfl_xor_sync(0xffffffffu, a.y, o);
    }
    G6[q] = cadd(a, to_cd(g0e[6 * (int64_t)j + q]));
  }
  cd G[3][3];
  G[0][0] = G6[0]; G[1][1] = G6[1]; G[2][2] = G6[2];
  G[0][1] = G[1][0] = G6[3];
  G[1][2] = G[2][1] = G6[4];
  G[2][0] = G[0][2] = G6[5];
  // ---- 3x3 system  (I + j s M G) y = j s M phi0[p]
  const double s = s_coef * area[j];
  const cd f[3] = {to_cd(ph[p0]), to_cd(ph[p1]), to_cd(ph[p2])};
  cd A3[3][3], r[3], y[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    cd mr = mk(0.0, 0.0);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const double mac = (a == c) ? (2.0 / 12.0) : (1.0 / 12.0);
      mr = cadd(mr, scale(mac, f[c]));
    }
    r[a] = mk(-s * mr.y, s * mr.x);   // j*s*mr
#pragma unroll
    for (int bb = 0; bb < 3; ++bb) {
      cd mg = mk(0.0, 0.0);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const double mac = (a == c) ? (2.0 / 12.0) : (1.0 / 12.0);
        mg = cadd(mg, scale(mac, G[c][bb]));
      }
      cd v = mk(-s * mg.y, s * mg.x);
      if (a == bb) v.x += 1.0;
      A3[a][bb] = v;
    }
  }
  solve3(A3, r, y);
  // ---- perturbation at the electrode nodes, amplitude
  const T* pu = ph + N0;
  for (int k = lane; k < nU; k += 32) {
    cd v = to_cd(pu[k]);
    v = fma_(to_cd(y0[k]), y[0], v);
    v = fma_(to_cd(y1[k]), y[1], v);
    v = fma_(to_cd(y2[k]), y[2], v);
    amp[k] = sqrt(fma(v.x, v.x, v.y * v.y));
  }
  __syncwarp();
  for (int m = lane; m < L; m += 32) {
    if (m == i) continue;
    double acc = 0.0;
    for (int q = we_ptr[m]; q < we_ptr[m + 1]; ++q) acc = fma(we_w[q], amp[we_u[q]], acc);
    J[((int64_t)i * (L - 1) + m - (m > i ? 1 : 0)) * ldJ + j] = fabs(acc);
  }
}

// ------------------------------------------------------------------------------------------------
// K2, patch form: one CTA per (element patch, excitation).
//  * The rows of Yh_i and Xh of the patch's (<= cap) distinct nodes are staged ONCE in shared memory
//    (coalesced; odd row pitch so that lanes reading different rows hit different banks).  Elements of a
//    patch share nodes: ~0.7 distinct nodes per element on a triangulated grid instead of 3.
//  * Elements are processed 32 at a time with ONE LANE PER ELEMENT, so the 3x3 complex solve and every
//    reduction are private to a lane (no shuffles, no redundant work); the 8 warps of the CTA split the
//    nU electrode nodes by electrode: each warp accumulates partial dot products over its electrodes'
//    nodes (combined through shared memory) and then produces the amplitudes and the weighted electrode
//    sums of exactly those electrodes, writing their J entries itself.
// Requires the electrode node ranges to be contiguous in U (no node shared by two electrodes).
// ------------------------------------------------------------------------------------------------
// |phi0 + d| = |phi0| sqrt(1 + x), x = (2 Re(conj(phi0) d) + |d|^2) / |phi0|^2.  For the finite-difference
// perturbations of this path |x| ~ 1e-8, so sqrt(1+x) - 1 is evaluated by its Taylor polynomial (error
// < 7/256 x^5 < 3e-17 for |x| <= 1e-3): 5 FMAs instead of a ~28-instruction DSQRT, and free of the
// cancellation in sqrt(1+x) - 1.  Larger |x| takes the exact branch.
__device__ __forceinline__ double sqrt1p_minus1(double x) {
  if (fabs(x) <= 1e-3) {
    double t = fma(x, -0.0390625, 0.0625);
    t = fma(x, t, -0.125);
    t = fma(x, t, 0.5);
    return x * t;
  }
  return sqrt(1.0 + x) - 1.0;
}
// The polynomial alone (no branch), for loops that check the range of x once per electrode: a loop body with the
// DSQRT branch is not if-converted, so its unrolled iterations serialise on convergence barriers.
__device__ __forceinline__ double sqrt1p_minus1_poly(double x) {
  double t = fma(x, -0.0390625, 0.0625);
  t = fma(x, t, -0.125);
  t = fma(x, t, 0.5);
  return x * t;
}
// high word of |x|: integer-pipe range tracking (hi >= SQRT1P_HI_LIMIT  <=>  |x| >= ~1e-3, conservative)
__device__ __forceinline__ unsigned abs_hi(double x) { return (unsigned)__double2hiint(x) & 0x7fffffffu; }
constexpr unsigned SQRT1P_HI_LIMIT = 0x3f50624du;      // high word of 1e-3

constexpr int WBP_WARPS = 8;

__device__ long long g_wb_clk[16];
#ifdef CEIT_TILE_CLOCKS
#define WB_CLK(q) do { if (threadIdx.x == 0 && (blockIdx.x == 81 || blockIdx.x == 1500)) g_wb_clk[(blockIdx.x == 81 ? 0 : 8) + q] = clock64(); } while (0)
#else
#define WB_CLK(q) do { } while (0)
#endif
// persistent kernel: per item of CTA 7, consumer group 0 (6 stamps) in [16 it, 16 it + 6) and the producer (3 stamps)
// in [16 it + 8, 16 it + 11)
__device__ long long g_wbq_clk[32 * 16];
#ifdef CEIT_TILE_CLOCKS
#define WBQ_CLK(it, q) do { if (blockIdx.x == 7 && lane == 0 && (it) < 32) g_wbq_clk[16 * (it) + (q)] = clock64(); } while (0)
#else
#define WBQ_CLK(it, q) do { } while (0)
#endif

// Real-base closed form of y = (I + j B)^-1 (j g), B = s M G real, g = s M f real (f = phi0 at the
// element's nodes, real on the real path):  (I + jB)^-1 = C (I - jB), C = (I + B^2)^-1  =>
// Re y = C B g,  Im y = C g.   ~120 real FMAs instead of a complex 3x3 adjugate.
__device__ __forceinline__ void solve3_real(const double G6[6], double s, const double f[3], double yr[3],
                                            double yi[3]) {
  // G (symmetric): [0]=00 [1]=11 [2]=22 [3]=01 [4]=12 [5]=20
  const double G[3][3] = {{G6[0], G6[3], G6[5]}, {G6[3], G6[1], G6[4]}, {G6[5], G6[4], G6[2]}};
  double B[3][3], g[3];
  const double s12 = s * (1.0 / 12.0);
#pragma unroll
  for (int b = 0; b < 3; ++b) {
    const double cs = G[0][b] + G[1][b] + G[2][b];
#pragma unroll
    for (int a = 0; a < 3; ++a) B[a][b] = s12 * (G[a][b] + cs);      // (M G)_ab = (G_ab + colsum_b) / 12
  }
  {
    const double fs = f[0] + f[1] + f[2];
#pragma unroll
    for (int a = 0; a < 3; ++a) g[a] = s12 * (f[a] + fs);
  }
  double A[3][3];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      double v = (a == b) ? 1.0 : 0.0;
#pragma unroll
      for (int c = 0; c < 3; ++c) v = fma(B[a][c], B[c][b], v);
      A[a][b] = v;
    }
  const double c00 = A[1][1] * A[2][2] - A[1][2] * A[2][1];
  const double c01 = A[1][2] * A[2][0] - A[1][0] * A[2][2];
  const double c02 = A[1][0] * A[2][1] - A[1][1] * A[2][0];
  const double det = A[0][0] * c00 + A[0][1] * c01 + A[0][2] * c02;
  const double id = 1.0 / det;
  const double c10 = A[0][2] * A[2][1] - A[0][1] * A[2][2];
  const double c11 = A[0][0] * A[2][2] - A[0][2] * A[2][0];
  const double c12 = A[0][1] * A[2][0] - A[0][0] * A[2][1];
  const double c20 = A[0][1] * A[1][2] - A[0][2] * A[1][1];
  const double c21 = A[0][2] * A[1][0] - A[0][0] * A[1][2];
  const double c22 = A[0][0] * A[1][1] - A[0][1] * A[1][0];
  double bg[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) bg[a] = B[a][0] * g[0] + B[a][1] * g[1] + B[a][2] * g[2];
  // C = adj(A) / det, adj[i][j] = cofactor[j][i]
  yi[0] = id * (c00 * g[0] + c10 * g[1] + c20 * g[2]);
  yi[1] = id * (c01 * g[0] + c11 * g[1] + c21 * g[2]);
  yi[2] = id * (c02 * g[0] + c12 * g[1] + c22 * g[2]);
  yr[0] = id * (c00 * bg[0] + c10 * bg[1] + c20 * bg[2]);
  yr[1] = id * (c01 * bg[0] + c11 * bg[1] + c21 * bg[2]);
  yr[2] = id * (c02 * bg[0] + c12 * bg[1] + c22 * bg[2]);
}

// Per-excitation constants of the electrode nodes for the real path: {2/phi0, 1/phi0^2, w |phi0|, 0}.
__global__ void node_consts_kernel(int nU, int64_t N, int64_t N0, const double* __restrict__ phi0,
                                   const double* __restrict__ wk, double4* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.y;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= nU) return;
  const double f = phi0[(int64_t)b * N + N0 + k];
  const double r = 1.0 / f;
  out[(int64_t)b * nU + k] = make_double4(2.0 * r, r * r, wk[k] * fabs(f), 0.0);
}

// Real path (zero base variable). One CTA per (patch of <= 32 elements, excitation); lane = element.
__global__ void __launch_bounds__(WBP_WARPS * 32)
woodbury_patch_kernel(int e_from, int e_to, int L, int nU, int pitch, int64_t N, int64_t N0,
                      const int32_t* __restrict__ patch_eptr, const int32_t* __restrict__ patch_elems,
                      const int32_t* __restrict__ patch_nptr, const int32_t* __restrict__ patch_nodes,
                      const uint8_t* __restrict__ patch_local, const double* __restrict__ area,
                      const double* __restrict__ g0e, const double* __restrict__ Xh,
                      const double* __restrict__ Yh, const double* __restrict__ phi0,
                      const int32_t* __restrict__ ue_ptr, const double4* __restrict__ node_consts, int i0,
                      double s_coef, double* __restrict__ J, int64_t ldJ, int cap, int nbatch) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int WARPS = WBP_WARPS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // excitation fastest: the CTAs that share a patch (same Xh rows) are scheduled together
  const int b = blockIdx.x % nbatch, patch = blockIdx.x / nbatch;
  const int i = i0 + b;
  const int n_beg = patch_nptr[patch], nn = patch_nptr[patch + 1] - n_beg;
  const int el_beg = patch_eptr[patch], ne = patch_eptr[patch + 1] - el_beg;
  double4* cU = reinterpret_cast<double4*>(smem_raw);           // [nU] {2/phi0, 1/phi0^2, w |phi0|, -}
  double* Ys = reinterpret_cast<double*>(cU + nU);              // [cap][pitch]
  double* Xs = Ys + (int64_t)cap * pitch;                       // [cap][pitch]
  double* part = Xs + (int64_t)cap * pitch;                     // [WARPS][6][32] partial dots
  double* ybuf = part + WARPS * 6 * 32;                         // [6][32] Re y0..2, Im y0..2
  double* phs = ybuf + 6 * 32;                                  // [cap] phi0 at the patch nodes
  const double* Y = Yh + (int64_t)b * N * nU;
  const double* ph = phi0 + (int64_t)b * N;
  WB_CLK(0);
  // ---- stage node rows with cp.async: every copy of the CTA is in flight at once -----------------
  for (int n = warp; n < nn; n += WARPS) {
    const int64_t p = patch_nodes[n_beg + n];
    const double* ys = Y + p * nU;
    const double* xs = Xh + p * nU;
    double* yd = Ys + (int64_t)n * pitch;
    double* xd = Xs + (int64_t)n * pitch;
    for (int k = lane; k < nU; k += 32) {
      cp_async8(yd + k, ys + k, true);
      cp_async8(xd + k, xs + k, true);
    }
    if (lane == 0) phs[n] = ph[p];
  }
  cp_async_commit();
  {
    const double4* cg = node_consts + (int64_t)b * nU;   // written once per excitation by node_consts_kernel
    for (int k = threadIdx.x; k < nU; k += WARPS * 32) cU[k] = cg[k];
  }
  // this lane's element
  const bool have = lane < ne;
  const int lq = have ? lane : 0;
  const int j = patch_elems[el_beg + lq];
  const int l0 = patch_local[3 * (el_beg + lq)], l1 = patch_local[3 * (el_beg + lq) + 1],
            l2 = patch_local[3 * (el_beg + lq) + 2];
  double g6[6];
  double sj = 0.0;
  if (warp == 0) {                                       // the solver warp prefetches its scalars
#pragma unroll
    for (int q = 0; q < 6; ++q) g6[q] = g0e[6 * (int64_t)j + q];
    sj = s_coef * area[j];
  }
  WB_CLK(1);
  cp_async_wait<0>();
  __syncthreads();
  WB_CLK(2);
  const double* y0 = Ys + (int64_t)l0 * pitch;
  const double* y1 = Ys + (int64_t)l1 * pitch;
  const double* y2 = Ys + (int64_t)l2 * pitch;
  const double* x0 = Xs + (int64_t)l0 * pitch;
  const double* x1 = Xs + (int64_t)l1 * pitch;
  const double* x2 = Xs + (int64_t)l2 * pitch;
  // ---- pass 1: partial dots over this warp's electrodes' nodes -----------------------------------
  // Lanes read DIFFERENT rows at the same k: the odd row pitch spreads the row bases over all 16
  // 8-byte bank pairs.
  double d[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  for (int m = warp; m < L; m += WARPS) {
    if (m == i) continue;                                // Dirichlet columns of Yh_i are zero
    const int ks = ue_ptr[m], ke = ue_ptr[m + 1];
#pragma unroll 4
    for (int k = ks; k < ke; ++k) {
      const double a0 = y0[k], a1 = y1[k], a2 = y2[k];
      const double b0 = x0[k], b1 = x1[k], b2 = x2[k];
      d[0] = fma(a0, b0, d[0]);
      d[1] = fma(a1, b1, d[1]);
      d[2] = fma(a2, b2, d[2]);
      d[3] = fma(a0, b1, d[3]);
      d[4] = fma(a1, b2, d[4]);
      d[5] = fma(a2, b0, d[5]);
    }
  }
#pragma unroll
  for (int q = 0; q < 6; ++q) part[(warp * 6 + q) * 32 + lane] = d[q];
  WB_CLK(3);
  __syncthreads();
  // ---- one warp combines the partial dots and solves the 3x3 systems of the 32 elements ----------
  if (warp == 0) {
#pragma unroll
    for (int q = 0; q < 6; ++q) {
      double acc = g6[q];
#pragma unroll
      for (int w = 0; w < WARPS; ++w) acc += part[(w * 6 + q) * 32 + lane];
      g6[q] = acc;
    }
    const double f[3] = {phs[l0], phs[l1], phs[l2]};
    double yr[3], yi[3];
    solve3_real(g6, sj, f, yr, yi);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      ybuf[a * 32 + lane] = yr[a];
      ybuf[(3 + a) * 32 + lane] = yi[a];
    }
  }
  WB_CLK(4);
  __syncthreads();
  WB_CLK(5);
  const double yr0 = ybuf[lane], yr1 = ybuf[32 + lane], yr2 = ybuf[64 + lane];
  const double yi0 = ybuf[96 + lane], yi1 = ybuf[128 + lane], yi2 = ybuf[160 + lane];
  // ---- pass 2: amplitudes and weighted sums of this warp's electrodes ----------------------------
  const bool write = have && j >= e_from && j < e_to;
  for (int m = warp; m < L; m += WARPS) {
    if (m == i) continue;
    const int ks = ue_ptr[m], ke = ue_ptr[m + 1];
    double sum0 = 0.0, sumt = 0.0;
    unsigned hmax = 0u;
#pragma unroll 4
    for (int k = ks; k < ke; ++k) {
      const double4 cu = cU[k];
      const double a0 = y0[k], a1 = y1[k], a2 = y2[k];
      const double dr = fma(a0, yr0, fma(a1, yr1, a2 * yr2));
      const double di = fma(a0, yi0, fma(a1, yi1, a2 * yi2));
      const double qq = fma(dr, dr, di * di);
      const double x = fma(cu.x, dr, qq * cu.y);         // (2 phi0 dr + |d|^2) / phi0^2, phi0 real
      hmax = max(hmax, abs_hi(x));
      sum0 += cu.z;
      sumt = fma(cu.z, sqrt1p_minus1_poly(x), sumt);
    }
    if (hmax >= SQRT1P_HI_LIMIT) {                       // rare: a large perturbation, exact square roots
      sumt = 0.0;
      for (int k = ks; k < ke; ++k) {
        const double4 cu = cU[k];
        const double a0 = y0[k], a1 = y1[k], a2 = y2[k];
        const double dr = fma(a0, yr0, fma(a1, yr1, a2 * yr2));
        const double di = fma(a0, yi0, fma(a1, yi1, a2 * yi2));
        const double qq = fma(dr, dr, di * di);
        sumt = fma(cu.z, sqrt1p_minus1(fma(cu.x, dr, qq * cu.y)), sumt);
      }
    }
    if (write) J[((int64_t)i * (L - 1) + m - (m > i ? 1 : 0)) * ldJ + j] = fabs(sum0 + sumt);
  }
  WB_CLK(6);
}

// ------------------------------------------------------------------------------------------------
// K2, low-rank form (rank-update excitation path).  With Yh_i = Y_T - Y_T[:,B] D_i + y_e (alpha q)^T the
// 3x3 block of Kff(i)^-1 on an element's nodes splits into an excitation-INDEPENDENT part and a
// rank-(b+1) correction:
//   G[a,b'] = g1e[pair] - sum_{j<b} Y_T[a,B_j] ZW_i[b'][j] + y_e[a] ZW_i[b'][bp],
//   g1e = G0 + (Y_T Xh^T)[a,b'] (once per material state, g1_kernel),  ZW_i = Xh Dcat_i^T (one batched GEMM).
// So the per-(excitation, element) dot products over all nU electrode nodes (6 nU FMAs and 6 nU shared loads)
// become 6 (b+1) FMAs, the Xh rows are no longer staged, and a patch needs half the shared memory.
// ------------------------------------------------------------------------------------------------
// g1e[e][q] = g0e[e][q] + sum_k Y_T[pa,k] Xh[pb,k]; one warp per element
__global__ void __launch_bounds__(256, 5)   // <= 51 registers: 40 warps per SM hold MESH_50_50's 5000 elements in one wave (was 1.06)
g1_kernel(int64_t E, int nU, const int32_t* __restrict__ elem_pos, const double* __restrict__ g0e,
          const double* __restrict__ YT, const double* __restrict__ Xh, double* __restrict__ g1e) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t e = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (e >= E) return;
  const int64_t p0 = elem_pos[3 * e], p1 = elem_pos[3 * e + 1], p2 = elem_pos[3 * e + 2];
  if ((p0 | p1 | p2) < 0) return;                      // decomposed plan: an element of another rank's region
  const double *y0 = YT + p0 * nU, *y1 = YT + p1 * nU, *y2 = YT + p2 * nU;
  const double *x0 = Xh + p0 * nU, *x1 = Xh + p1 * nU, *x2 = Xh + p2 * nU;
  double d[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  for (int k = lane; k < nU; k += 32) {
    const double a0 = y0[k], a1 = y1[k], a2 = y2[k];
    const double b0 = x0[k], b1 = x1[k], b2 = x2[k];
    d[0] = fma(a0, b0, d[0]);
    d[1] = fma(a1, b1, d[1]);
    d[2] = fma(a2, b2, d[2]);
    d[3] = fma(a0, b1, d[3]);
    d[4] = fma(a1, b2, d[4]);
    d[5] = fma(a2, b0, d[5]);
  }
#pragma unroll
  for (int q = 0; q < 6; ++q) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) d[q] += __shfl_xor_sync(0xffffffffu, d[q], o);
  }
  if (lane < 6) {
    double v = d[0];
#pragma unroll
    for (int q = 1; q < 6; ++q) v = (lane == q) ? d[q] : v;
    g1e[6 * e + lane] = g0e[6 * e + lane] + v;
  }
}

// ---- TMA bulk copy (1-D cp.async.bulk, SASS UBLKCP) + mbarrier helpers --------------------------------------------
// One instruction moves a whole node row (nU * 8 bytes) global -> shared; an 8-byte LDGSTS costs ~8 LSU cycles per
// warp instruction whatever its size, which made the staging loop of this kernel 30 % of a CTA's life.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// Per-excitation constants of the low-rank kernel, structure of arrays so that ONE bulk copy stages them and the
// main loop reads two electrode nodes per 16-byte broadcast load:
//   out[b] = [ 2/phi0 (nUp) | 1/phi0^2 (nUp) | w |phi0| (nUp) | sum over electrode m of w |phi0| (Lp) ]
__global__ void node_consts_soa_kernel(int nU, int nUp, int L, int cst_len, int64_t N, int64_t N0,
                                       const double* __restrict__ phi0, const double* __restrict__ wk,
                                       const int32_t* __restrict__ ue_ptr, double* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.y;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  const double* ph = phi0 + (int64_t)b * N + N0;
  double* o = out + (int64_t)b * cst_len;
  if (k < nU) {
    const double f = ph[k];
    const double r = 1.0 / f;
    o[k] = 2.0 * r;
    o[nUp + k] = r * r;
    o[2 * nUp + k] = wk[k] * fabs(f);
  }
  if (k < L) {
    double s0 = 0.0;
    for (int q = ue_ptr[k]; q < ue_ptr[k + 1]; ++q) s0 += wk[q] * fabs(ph[q]);
    o[3 * nUp + k] = s0;
  }
}

// Row pitch of the staged Yh rows: even (16-byte rows for the bulk copy and the 16-byte loads) with pitch/2 odd, so
// that the 8 lanes of a quarter warp reading different rows at the same column pair hit 8 different 16-byte bank groups.
__host__ __device__ inline int woodbury_lr_pitch(int nU) {
  int p = (nU + 1) & ~1;
  if (((p >> 1) & 1) == 0) p += 2;
  return p;
}

__global__ void __launch_bounds__(WBP_WARPS * 32)
woodbury_lr_kernel(int e_from, int e_to, int L, int nU, int pitch, int bp, int64_t N, int64_t N0,
                   const int32_t* __restrict__ patch_eptr, const int32_t* __restrict__ patch_elems,
                   const int32_t* __restrict__ patch_nptr, const int32_t* __restrict__ patch_nodes,
                   const uint8_t* __restrict__ patch_local, const double* __restrict__ area,
                   const double* __restrict__ g1e, const double* __restrict__ YT, const double* __restrict__ ye,
                   const double* __restrict__ ZW, const double* __restrict__ Yh, const double* __restrict__ phi0,
                   const int32_t* __restrict__ ue_ptr, const double* __restrict__ node_consts, int i0,
                   double s_coef, double* __restrict__ J, int64_t ldJ, int cap, int nbatch, int64_t ldz) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int WARPS = WBP_WARPS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x % nbatch, patch = blockIdx.x / nbatch;   // excitation fastest
  const int i = i0 + b;
  const int n_beg = patch_nptr[patch], nn = patch_nptr[patch + 1] - n_beg;
  const int el_beg = patch_eptr[patch], ne = patch_eptr[patch + 1] - el_beg;
  const int lo = ue_ptr[i], nb = ue_ptr[i + 1] - lo;
  const int zp = bp + 1;                                         // odd stride of the small per-node arrays
  const int nUp = (nU + 1) & ~1, Lp = (L + 1) & ~1, cst_len = 3 * nUp + Lp;
  const bool bulk = (nU & 1) == 0;                               // 16-byte aligned rows: TMA bulk copies
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);         // 16 bytes reserved
  double* cst = reinterpret_cast<double*>(smem_raw + 16);        // [3 nUp + Lp]  per-excitation constants (SoA)
  double* Ys = cst + cst_len;                                    // [cap][pitch]  rows of Yh_i
  double* YB = Ys + (int64_t)cap * pitch;                        // [cap][zp]     Y_T[node, B_i]
  double* ZWs = YB + cap * zp;                                   // [cap][zp]     ZW_i[node, :]
  double* yes = ZWs + cap * zp;                                  // [cap]         y_e[node]
  double* phs = yes + cap;                                       // [cap]         phi0[node]
  double* ybuf = phs + cap;                                      // [6][32]       Re y0..2, Im y0..2
  double* gbuf = ybuf + 6 * 32;                                  // [6][32]       corrected 3x3 blocks
  const double* Y = Yh + (int64_t)b * N * nU;
  const double* ph = phi0 + (int64_t)b * N;
  const double* Z = ZW + (int64_t)b * zp;                        // ZW is N x (nbatch zp), excitation b's columns
  const double* cg = node_consts + (int64_t)b * cst_len;
  WB_CLK(0);
  if (threadIdx.x == 0 && bulk) mbar_init(bar, 1);
  // node positions of the patch: one coalesced load
  const int pn_a = (lane < nn) ? patch_nodes[n_beg + lane] : 0;
  const int pn_b = (lane + 32 < nn) ? patch_nodes[n_beg + lane + 32] : 0;
  __syncthreads();                                               // barrier initialised before anyone polls it
  WB_CLK(1);
  if (bulk) {
    // warp 0: one bulk copy per node row of Yh_i (lane = node) + the constants; everything lands on `bar`
    if (warp == 0) {
      if (lane == 0) {
        mbar_expect_tx(bar, (uint32_t)(((int64_t)nn * nU + cst_len) * 8));
        bulk_g2s(cst, cg, (uint32_t)(cst_len * 8), bar);
      }
      __syncwarp();
      if (lane < nn) bulk_g2s(Ys + (int64_t)lane * pitch, Y + (int64_t)pn_a * nU, (uint32_t)(nU * 8), bar);
      if (lane + 32 < nn) bulk_g2s(Ys + (int64_t)(lane + 32) * pitch, Y + (int64_t)pn_b * nU, (uint32_t)(nU * 8), bar);
    }
  } else {
    for (int n = warp; n < nn; n += WARPS) {
      const int64_t p = __shfl_sync(0xffffffffu, (n < 32) ? pn_a : pn_b, n & 31);
      for (int k = lane; k < nU; k += 32) cp_async8(Ys + (int64_t)n * pitch + k, Y + p * nU + k, true);
    }
    for (int k = threadIdx.x; k < cst_len; k += WARPS * 32) cp_async8(cst + k, cg + k, true);
  }
  // the small per-node arrays (unaligned sources): 8-byte cp.async, a few instructions per warp
  for (int n = warp; n < nn; n += WARPS) {
    const int64_t p = __shfl_sync(0xffffffffu, (n < 32) ? pn_a : pn_b, n & 31);
    for (int k = lane; k < nb; k += 32) cp_async8(YB + n * zp + k, YT + p * nU + lo + k, true);
    for (int k = lane; k < zp; k += 32) cp_async8(ZWs + n * zp + k, Z + p * ldz + k, true);
    if (lane == 0) cp_async8(yes + n, ye + p, true);
    if (lane == 1) cp_async8(phs + n, ph + p, true);
  }
  cp_async_commit();
  WB_CLK(2);
  const bool have = lane < ne;
  const int lq = have ? lane : 0;
  const int j = patch_elems[el_beg + lq];
  const int l0 = patch_local[3 * (el_beg + lq)], l1 = patch_local[3 * (el_beg + lq) + 1],
            l2 = patch_local[3 * (el_beg + lq) + 2];
  // warps 0..5 each correct one entry of the symmetric 3x3 block (pairs (0,0),(1,1),(2,2),(0,1),(1,2),(2,0))
  // for the 32 elements; warp 0 then solves
  const int la = (warp == 0 || warp == 3) ? l0 : ((warp == 1 || warp == 4) ? l1 : l2);
  const int lb = (warp == 0 || warp == 5) ? l0 : ((warp == 1 || warp == 3) ? l1 : l2);
  double gq = 0.0, sj = 0.0;
  if (warp < 6) gq = g1e[6 * (int64_t)j + warp];
  if (warp == 0) sj = s_coef * area[j];
  WB_CLK(3);
  cp_async_wait<0>();
  __syncthreads();                                               // the small arrays of every warp are visible
  WB_CLK(4);
  if (warp < 6) {
    const double* yb = YB + la * zp;
    const double* zw = ZWs + lb * zp;
    double a0 = yes[la] * zw[bp], a1 = 0.0;
    int jj = 0;
    for (; jj + 1 < nb; jj += 2) {
      a0 = fma(-yb[jj], zw[jj], a0);
      a1 = fma(-yb[jj + 1], zw[jj + 1], a1);
    }
    if (jj < nb) a0 = fma(-yb[jj], zw[jj], a0);
    gbuf[warp * 32 + lane] = gq + (a0 + a1);
  }
  __syncthreads();
  WB_CLK(5);
  // ---- one warp: the 3x3 solve, one lane per element -----------------------------------------------------
  if (warp == 0) {
    double g6[6];
#pragma unroll
    for (int q = 0; q < 6; ++q) g6[q] = gbuf[q * 32 + lane];
    const double f[3] = {phs[l0], phs[l1], phs[l2]};
    double yr[3], yi[3];
    solve3_real(g6, sj, f, yr, yi);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      ybuf[a * 32 + lane] = yr[a];
      ybuf[(3 + a) * 32 + lane] = yi[a];
    }
  }
  if (bulk) mbar_wait(bar, 0);                                   // the Yh rows and constants (async proxy) have landed
  __syncthreads();
  WB_CLK(6);
  const double yr0 = ybuf[lane], yr1 = ybuf[32 + lane], yr2 = ybuf[64 + lane];
  const double yi0 = ybuf[96 + lane], yi1 = ybuf[128 + lane], yi2 = ybuf[160 + lane];
  const double* y0 = Ys + (int64_t)l0 * pitch;
  const double* y1 = Ys + (int64_t)l1 * pitch;
  const double* y2 = Ys + (int64_t)l2 * pitch;
  const double *cx = cst, *cy = cst + nUp, *cz = cst + 2 * nUp, *s0 = cst + 3 * nUp;
  // ---- amplitudes and weighted sums of this warp's electrodes -------------------------------------------
  // term(k) = w |phi0| (sqrt(1 + x) - 1), x = (2 phi0 dr + |d|^2) / phi0^2 (phi0 real); two nodes per 16-byte load
#define CEIT_WB_TERM(ACC, A0, A1, A2, CX, CY, CZ)                 \
  {                                                               \
    const double dr = fma(A0, yr0, fma(A1, yr1, (A2) * yr2));     \
    const double di = fma(A0, yi0, fma(A1, yi1, (A2) * yi2));     \
    const double qq = fma(dr, dr, di * di);                       \
    const double x = fma(CX, dr, qq * (CY));                      \
    hmax = max(hmax, abs_hi(x));                                  \
    ACC = fma(CZ, sqrt1p_minus1_poly(x), ACC);                    \
  }
  const bool write = have && j >= e_from && j < e_to;
  for (int m = warp; m < L; m += WARPS) {
    if (m == i) continue;
    const int ks = ue_ptr[m], ke = ue_ptr[m + 1];
    double sumt = 0.0, sumu = 0.0;
    unsigned hmax = 0u;
    int k = ks;
    if ((k & 1) && k < ke) {
      CEIT_WB_TERM(sumt, y0[k], y1[k], y2[k], cx[k], cy[k], cz[k]);
      ++k;
    }
#pragma unroll 2
    for (; k + 1 < ke; k += 2) {
      const double2 a0 = *reinterpret_cast<const double2*>(y0 + k);
      const double2 a1 = *reinterpret_cast<const double2*>(y1 + k);
      const double2 a2 = *reinterpret_cast<const double2*>(y2 + k);
      const double2 vx = *reinterpret_cast<const double2*>(cx + k);
      const double2 vy = *reinterpret_cast<const double2*>(cy + k);
      const double2 vz = *reinterpret_cast<const double2*>(cz + k);
      CEIT_WB_TERM(sumt, a0.x, a1.x, a2.x, vx.x, vy.x, vz.x);
      CEIT_WB_TERM(sumu, a0.y, a1.y, a2.y, vx.y, vy.y, vz.y);
    }
    if (k < ke) CEIT_WB_TERM(sumt, y0[k], y1[k], y2[k], cx[k], cy[k], cz[k]);
    sumt += sumu;
    if (hmax >= SQRT1P_HI_LIMIT) {                       // rare: a large perturbation, exact square roots
      sumt = 0.0;
      for (int q = ks; q < ke; ++q) {
        const double dr = fma(y0[q], yr0, fma(y1[q], yr1, y2[q] * yr2));
        const double di = fma(y0[q], yi0, fma(y1[q], yi1, y2[q] * yi2));
        const double qq = fma(dr, dr, di * di);
        sumt = fma(cz[q], sqrt1p_minus1(fma(cx[q], dr, qq * cy[q])), sumt);
      }
    }
    if (write) J[((int64_t)i * (L - 1) + m - (m > i ? 1 : 0)) * ldJ + j] = fabs(s0[m] + sumt);
  }
#undef CEIT_WB_TERM
  WB_CLK(7);
}

// ------------------------------------------------------------------------------------------------
// K2, persistent pipelined form of the low-rank kernel (default when the shapes allow it).
// The one-CTA-per-(patch, excitation) kernel above spends 45 % of a CTA's life before its main loop (staging,
// 3x3 correction and solve; tools/wb_clocks.py) and a third of the SM time between CTAs.  Here ONE CTA per SM stays
// resident and walks its share of the work items
//     item = (patch, excitation, column chunk)         (chunks = electrode ranges whose first column is even)
// through a ring of WBQ_SLOTS shared-memory slots: WBQ_PW PRODUCER warps stage item k + 2 (TMA bulk copies of the patch's
// Yh rows and of the per-excitation constants, 8-byte cp.async for the small per-node arrays, all completing on the
// slot's `full` mbarrier) while WBQ_GROUPS consumer groups of 8 warps each work on items k and k + 1 (correction of
// the 3x3 blocks, solve, amplitude loop - the code of the kernel above); a group releases its slot through the `empty`
// mbarrier.  Launch overhead, staging latency and the solve of one item hide behind the amplitude loops of the others.
// Requires even nU (16-byte aligned rows).  Shared memory: 64 B of barriers, 12 x 32 doubles per group, then the slots
//   [ cx | cy | cz (Wp each) | s0 (Lp) | Ys cap x pitchW | YB cap x zp | ZW cap x zp | ye cap | phi0 cap ].
// ------------------------------------------------------------------------------------------------
constexpr int WBQ_GROUPS = 2, WBQ_SLOTS = 3, WBQ_GW = 8;          // consumer groups, ring slots, warps per group
constexpr int WBQ_PW = 4;   // producer warps: a bulk copy costs ~75 issue cycles and the copies of one warp serialise
constexpr int WBQ_THREADS = (WBQ_GROUPS * WBQ_GW + WBQ_PW) * 32;
struct WbqChunks {
  int n;            // column chunks per (patch, excitation)
  int el[5];        // electrode boundaries of the chunks: chunk c = electrodes [el[c], el[c + 1])
  int wmax;         // widest chunk in columns (even)
};
// per (patch, lane): what a consumer lane needs of its element, in ONE 16-byte load that does not depend on
// patch_eptr (the chain patch_eptr -> patch_elems -> patch_local -> area was ~1.2k cycles per item)
struct __align__(16) WbqPatchRec {
  int32_t j;        // element id (lane >= elements of the patch: the first element's data, valid = 0)
  uint32_t pack;    // l0 | l1 << 8 | l2 << 16 | valid << 24
  double area;
};
__global__ void patch_rec_kernel(int n_patches, const int32_t* __restrict__ patch_eptr,
                                 const int32_t* __restrict__ patch_elems, const uint8_t* __restrict__ patch_local,
                                 const double* __restrict__ area, WbqPatchRec* __restrict__ rec) {
  pdl_trigger();
  pdl_wait();
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= (int64_t)n_patches * 32) return;
  const int patch = (int)(t >> 5), lane = (int)(t & 31);
  const int el_beg = patch_eptr[patch], ne = patch_eptr[patch + 1] - el_beg;
  const bool have = lane < ne;
  const int q = el_beg + (have ? lane : 0);
  WbqPatchRec r;
  r.j = patch_elems[q];
  r.pack = (uint32_t)patch_local[3 * q] | ((uint32_t)patch_local[3 * q + 1] << 8) | ((uint32_t)patch_local[3 * q + 2] << 16) |
           ((have ? 1u : 0u) << 24);
  r.area = area[r.j];
  rec[t] = r;
}
// slot: [ cx | cy | cz (Wp each) | s0 (Lp) | Ys cap x pitchW | YB cap x zs | ZW cap x zs | ye cap | phi0 cap ], zs = bp + 2
__host__ __device__ inline int64_t wbq_slot_doubles(int wmax, int L, int cap, int bp) {
  const int Wp = (wmax + 1) & ~1, Lp = (L + 1) & ~1;
  int64_t d = 3 * (int64_t)Wp + Lp + (int64_t)cap * woodbury_lr_pitch(wmax) + 2 * (int64_t)cap * (bp + 2) + 2 * cap;
  return (d + 1) & ~(int64_t)1;
}
__host__ __device__ inline int64_t wbq_smem_bytes(int wmax, int L, int cap, int bp) {
  return 64 + (int64_t)WBQ_GROUPS * 12 * 32 * 8 + WBQ_SLOTS * wbq_slot_doubles(wmax, L, cap, bp) * 8;
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
// the barrier receives one arrival when every cp.async issued so far by this thread has landed (counted in the
// barrier's initial count: .noinc)
__device__ __forceinline__ void cp_async_mbar_arrive(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void group_bar(int g) {
  asm volatile("bar.sync %0, %1;\n" ::"r"(1 + g), "r"(WBQ_GW * 32) : "memory");
}

__global__ void __launch_bounds__(WBQ_THREADS, 1)
woodbury_lr_persist_kernel(int e_from, int e_to, int L, int nU, int bp, int64_t N, int64_t N0, int n_items,
                           WbqChunks ch, const int32_t* __restrict__ patch_eptr,
                           const int32_t* __restrict__ patch_nptr, const int32_t* __restrict__ patch_nodes,
                           const WbqPatchRec* __restrict__ patch_rec, const double* __restrict__ g1e,
                           const double* __restrict__ YT, const double* __restrict__ ye,
                           const double* __restrict__ ZW, const double* __restrict__ Yh,
                           const double* __restrict__ phi0, const int32_t* __restrict__ ue_ptr,
                           const double* __restrict__ node_consts, int i0, double s_coef, double* __restrict__ J,
                           int64_t ldJ, int cap, int nbatch, int64_t ldz) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int zp = bp + 1, zs = bp + 2;                              // small rows: zp values inside aligned windows of zs
  const int nUp = (nU + 1) & ~1, Lp = (L + 1) & ~1, cst_len = 3 * nUp + Lp;
  const int Wp = (ch.wmax + 1) & ~1, pitch = woodbury_lr_pitch(ch.wmax);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw);          // [WBQ_SLOTS]
  uint64_t* empty = full + WBQ_SLOTS;                              // [WBQ_SLOTS]
  double* gbase = reinterpret_cast<double*>(smem_raw + 64);        // per group: ybuf [6][32], gbuf [6][32]
  double* slots = gbase + WBQ_GROUPS * 12 * 32;
  const int64_t slot_d = wbq_slot_doubles(ch.wmax, L, cap, bp);
  // contiguous share of the items (excitation fastest, then chunk, then patch)
  const int q0 = (int)(((int64_t)blockIdx.x * n_items) / gridDim.x);
  const int q1 = (int)(((int64_t)(blockIdx.x + 1) * n_items) / gridDim.x);
  const int n_my = q1 - q0;
  if (threadIdx.x == 0) {
    for (int sl = 0; sl < WBQ_SLOTS; ++sl) {
      mbar_init(full + sl, 1 + 32 * WBQ_PW);                       // expect_tx arrival + one cp.async arrival per producer thread
      mbar_init(empty + sl, WBQ_GW);                               // one arrival per consumer warp
    }
  }
  __syncthreads();
  const int per = nbatch * ch.n;
  if (warp >= WBQ_GROUPS * WBQ_GW) {
    // ---------------------------------------- producers -----------------------------------------------
    const int pw = warp - WBQ_GROUPS * WBQ_GW;                     // node n is staged by producer warp n % WBQ_PW
    // every array of an item is staged with bulk copies, one instruction per node row (lane = node): the rows of
    // Yh_i, and 16-byte aligned windows of zs doubles around Y_T[node, B_i] and ZW_i[node, :] (an 8-byte cp.async
    // costs ~80 issue cycles: 130 of them per item made one warp the bottleneck of the whole ring)
    for (int it = 0; it < n_my; ++it) {
      const int q = q0 + it, sl = it % WBQ_SLOTS, use = it / WBQ_SLOTS;
      const int patch = q / per, r = q - patch * per, c = r / nbatch, b = r - c * nbatch;
      const int i = i0 + b;
      WBQ_CLK(it, 8);
      const int n_beg = patch_nptr[patch], nn = patch_nptr[patch + 1] - n_beg;
      const int lo = ue_ptr[i];
      const int k0 = ue_ptr[ch.el[c]], W = ue_ptr[ch.el[c + 1]] - k0;
      const int pn_a = (lane < nn) ? patch_nodes[n_beg + lane] : 0;
      const int pn_b = (lane + 32 < nn) ? patch_nodes[n_beg + lane + 32] : 0;
      double* cst = slots + sl * slot_d;
      double* Ys = cst + 3 * Wp + Lp;
      double* YB = Ys + (int64_t)cap * pitch;
      double* ZWs = YB + cap * zs;
      double* yes = ZWs + cap * zs;
      double* phs = yes + cap;
      const double* Y = Yh + (int64_t)b * N * nU + k0;
      const double* ph = phi0 + (int64_t)b * N;
      const int lo_al = lo & ~1, len_y = min(zs, nU - lo_al);      // window of Y_T[node, B_i]
      const int z_al = (b * zp) & ~1;                              // window of ZW_i[node, :]
      const double* cg = node_consts + (int64_t)b * cst_len;
      if (use > 0) mbar_wait(empty + sl, (use - 1) & 1);           // the slot's previous item has been consumed
      WBQ_CLK(it, 9);
      if (lane == 0 && pw == 0) {
        mbar_expect_tx(full + sl, (uint32_t)(((int64_t)nn * (W + len_y + zs) + 3 * W + Lp) * 8));
        bulk_g2s(cst, cg + k0, (uint32_t)(W * 8), full + sl);
        bulk_g2s(cst + Wp, cg + nUp + k0, (uint32_t)(W * 8), full + sl);
        bulk_g2s(cst + 2 * Wp, cg + 2 * nUp + k0, (uint32_t)(W * 8), full + sl);
        bulk_g2s(cst + 3 * Wp, cg + 3 * nUp, (uint32_t)(Lp * 8), full + sl);
      }
      __syncwarp();
      if (lane < nn && (lane % WBQ_PW) == pw) {
        bulk_g2s(Ys + (int64_t)lane * pitch, Y + (int64_t)pn_a * nU, (uint32_t)(W * 8), full + sl);
        bulk_g2s(YB + lane * zs, YT + (int64_t)pn_a * nU + lo_al, (uint32_t)(len_y * 8), full + sl);
        bulk_g2s(ZWs + lane * zs, ZW + (int64_t)pn_a * ldz + z_al, (uint32_t)(zs * 8), full + sl);
        cp_async8(yes + lane, ye + pn_a, true);
        cp_async8(phs + lane, ph + pn_a, true);
      }
      if (lane + 32 < nn && (lane % WBQ_PW) == pw) {
        bulk_g2s(Ys + (int64_t)(lane + 32) * pitch, Y + (int64_t)pn_b * nU, (uint32_t)(W * 8), full + sl);
        bulk_g2s(YB + (lane + 32) * zs, YT + (int64_t)pn_b * nU + lo_al, (uint32_t)(len_y * 8), full + sl);
        bulk_g2s(ZWs + (lane + 32) * zs, ZW + (int64_t)pn_b * ldz + z_al, (uint32_t)(zs * 8), full + sl);
        cp_async8(yes + lane + 32, ye + pn_b, true);
        cp_async8(phs + lane + 32, ph + pn_b, true);
      }
      cp_async_mbar_arrive(full + sl);
      WBQ_CLK(it, 10);
    }
    return;
  }
  // ------------------------------------------ consumers ----------------------------------------------
  const int g = warp / WBQ_GW, wg = warp % WBQ_GW;
  double* ybuf = gbase + g * 12 * 32;
  double* gbuf = ybuf + 6 * 32;
  WbqPatchRec rec{};
  double gq = 0.0;
  if (g < n_my) {                                                  // element data of the first item of this group
    const int patch = (q0 + g) / per;
    rec = patch_rec[(int64_t)patch * 32 + lane];
    if (wg < 6) gq = g1e[6 * (int64_t)rec.j + wg];
  }
  for (int it = g; it < n_my; it += WBQ_GROUPS) {
    const int q = q0 + it, sl = it % WBQ_SLOTS, use = it / WBQ_SLOTS;
    const int patch = q / per, r = q - patch * per, c = r / nbatch, b = r - c * nbatch;
    const int i = i0 + b;
    if (warp == 0) WBQ_CLK(it, 0);
    const int lo = ue_ptr[i], nb = ue_ptr[i + 1] - lo;
    const int k0 = ue_ptr[ch.el[c]];
    const double* cst = slots + sl * slot_d;
    const double* Ys = cst + 3 * Wp + Lp;
    const double* YB = Ys + (int64_t)cap * pitch;
    const double* ZWs = YB + cap * zs;
    const double* yes = ZWs + cap * zs;
    const double* phs = yes + cap;
    const bool have = (rec.pack >> 24) != 0;
    const int j = rec.j;
    const int l0 = rec.pack & 255, l1 = (rec.pack >> 8) & 255, l2 = (rec.pack >> 16) & 255;
    const int la = (wg == 0 || wg == 3) ? l0 : ((wg == 1 || wg == 4) ? l1 : l2);
    const int lb = (wg == 0 || wg == 5) ? l0 : ((wg == 1 || wg == 3) ? l1 : l2);
    const double sj = s_coef * rec.area;
    const double gq_cur = gq;
    if (warp == 0) WBQ_CLK(it, 1);
    mbar_wait(full + sl, use & 1);
    if (warp == 0) WBQ_CLK(it, 2);
    if (wg < 6) {
      const double* yb = YB + la * zs + (lo & 1);
      const double* zw = ZWs + lb * zs + ((b * zp) & 1);
      double a0 = yes[la] * zw[bp], a1 = 0.0, a2 = 0.0, a3 = 0.0;
      int jj = 0;
      for (; jj + 3 < nb; jj += 4) {
        a0 = fma(-yb[jj], zw[jj], a0);
        a1 = fma(-yb[jj + 1], zw[jj + 1], a1);
        a2 = fma(-yb[jj + 2], zw[jj + 2], a2);
        a3 = fma(-yb[jj + 3], zw[jj + 3], a3);
      }
      for (; jj < nb; ++jj) a0 = fma(-yb[jj], zw[jj], a0);
      gbuf[wg * 32 + lane] = gq_cur + ((a0 + a1) + (a2 + a3));
    }
    group_bar(g);
    if (warp == 0) WBQ_CLK(it, 3);
    if (wg == 0) {
      double g6[6];
#pragma unroll
      for (int qq = 0; qq < 6; ++qq) g6[qq] = gbuf[qq * 32 + lane];
      const double f[3] = {phs[l0], phs[l1], phs[l2]};
      double yr[3], yi[3];
      solve3_real(g6, sj, f, yr, yi);
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        ybuf[a * 32 + lane] = yr[a];
        ybuf[(3 + a) * 32 + lane] = yi[a];
      }
    }
    // element data of this group's NEXT item: the loads fly during the amplitude loop
    WbqPatchRec rec_n = rec;
    const bool new_patch = it + WBQ_GROUPS < n_my && (q + WBQ_GROUPS) / per != patch;
    if (new_patch) rec_n = patch_rec[(int64_t)((q + WBQ_GROUPS) / per) * 32 + lane];
    group_bar(g);
    if (warp == 0) WBQ_CLK(it, 4);
    const double yr0 = ybuf[lane], yr1 = ybuf[32 + lane], yr2 = ybuf[64 + lane];
    const double yi0 = ybuf[96 + lane], yi1 = ybuf[128 + lane], yi2 = ybuf[160 + lane];
    // pointers biased by -k0 so that k stays the global electrode-node index (k0 is even: 16-byte alignment of even k)
    const double* y0 = Ys + (int64_t)l0 * pitch - k0;
    const double* y1 = Ys + (int64_t)l1 * pitch - k0;
    const double* y2 = Ys + (int64_t)l2 * pitch - k0;
    const double *cx = cst - k0, *cy = cst + Wp - k0, *cz = cst + 2 * Wp - k0, *s0 = cst + 3 * Wp;
#define CEIT_WB_TERM(ACC, A0, A1, A2, CX, CY, CZ)                 \
  {                                                               \
    const double dr = fma(A0, yr0, fma(A1, yr1, (A2) * yr2));     \
    const double di = fma(A0, yi0, fma(A1, yi1, (A2) * yi2));     \
    const double qq = fma(dr, dr, di * di);                       \
    const double x = fma(CX, dr, qq * (CY));                      \
    hmax = max(hmax, abs_hi(x));                                  \
    ACC = fma(CZ, sqrt1p_minus1_poly(x), ACC);                    \
  }
    const bool write = have && j >= e_from && j < e_to;
    for (int m = ch.el[c] + wg; m < ch.el[c + 1]; m += WBQ_GW) {
      if (m == i) continue;
      const int ks = ue_ptr[m], ke = ue_ptr[m + 1];
      double sumt = 0.0, sumu = 0.0;
      unsigned hmax = 0u;
      int k = ks;
      if ((k & 1) && k < ke) {
        CEIT_WB_TERM(sumt, y0[k], y1[k], y2[k], cx[k], cy[k], cz[k]);
        ++k;
      }
#pragma unroll 4
      for (; k + 1 < ke; k += 2) {
        const double2 a0 = *reinterpret_cast<const double2*>(y0 + k);
        const double2 a1 = *reinterpret_cast<const double2*>(y1 + k);
        const double2 a2 = *reinterpret_cast<const double2*>(y2 + k);
        const double2 vx = *reinterpret_cast<const double2*>(cx + k);
        const double2 vy = *reinterpret_cast<const double2*>(cy + k);
        const double2 vz = *reinterpret_cast<const double2*>(cz + k);
        CEIT_WB_TERM(sumt, a0.x, a1.x, a2.x, vx.x, vy.x, vz.x);
        CEIT_WB_TERM(sumu, a0.y, a1.y, a2.y, vx.y, vy.y, vz.y);
      }
      if (k < ke) CEIT_WB_TERM(sumt, y0[k], y1[k], y2[k], cx[k], cy[k], cz[k]);
      sumt += sumu;
      if (hmax >= SQRT1P_HI_LIMIT) {                       // rare: a large perturbation, exact square roots
        sumt = 0.0;
        for (int qx = ks; qx < ke; ++qx) {
          const double dr = fma(y0[qx], yr0, fma(y1[qx], yr1, y2[qx] * yr2));
          const double di = fma(y0[qx], yi0, fma(y1[qx], yi1, y2[qx] * yi2));
          const double qq = fma(dr, dr, di * di);
          sumt = fma(cz[qx], sqrt1p_minus1(fma(cx[qx], dr, qq * cy[qx])), sumt);
        }
      }
      if (write) J[((int64_t)i * (L - 1) + m - (m > i ? 1 : 0)) * ldJ + j] = fabs(s0[m] + sumt);
    }
#undef CEIT_WB_TERM
    __syncwarp();
    if (warp == 0) WBQ_CLK(it, 5);
    if (lane == 0) mbar_arrive(empty + sl);                 // this warp no longer reads the slot
    if (new_patch && wg < 6) gq = g1e[6 * (int64_t)rec_n.j + wg];   // lands during the next item's correction loop
    rec = rec_n;
  }
}

// ------------------------------------------------------------------------------------------------
// inverse model / realtime helpers
// ------------------------------------------------------------------------------------------------
// Jt[r][c] = J[rows[r]][det[c]] - shift           (Solver.py:76-86,119)
__global__ void gather_shift_kernel(int64_t m_sel, int64_t n_det, const double* __restrict__ J, int64_t ldJ,
                                    const int64_t* __restrict__ rows, const int64_t* __restrict__ det,
                                    double shift, double* __restrict__ Jt) {
  pdl_trigger();
  pdl_wait();
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= m_sel * n_det) return;
  const int64_t r = q / n_det, c = q - r * n_det;
  const int64_t rr = rows ? rows[r] : r;
  Jt[q] = J[rr * ldJ + det[c]] - shift;
}
__global__ void add_diag_kernel(int64_t n, double v, double* __restrict__ A, int64_t lda) {
  pdl_trigger();
  pdl_wait();
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q < n) A[q * lda + q] += v;
}
// out[r] = sum_k inv[r][k] dV[k]; one warp per row (single-frame Solver.solve)
__global__ void __launch_bounds__(256)
gemv_rows_kernel(int64_t n_det, int m, const double* __restrict__ inv, const double* __restrict__ dV,
                 double* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n_det) return;
  const double* r = inv + row * m;
  double acc = 0.0;
  for (int k = lane; k < m; k += 32) acc = fma(r[k], dV[k], acc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) out[row] = acc;
}

// Centre of position of reconstructed maps (DataHandler.get_COP, CEIT/datahandler/DataHandler.py:28-56), batched:
// maps[n_det][F] row-major (the layout ceit_reconstruct writes).  Per frame: threshold = frac max + (1 - frac) min,
// then the mean centroid (cx, cy) and mean value of the detection elements strictly above it.
// CTA = 32 consecutive frames (one per lane: coalesced row reads) x 8 warps over the rows; the warps' partial
// results are combined in warp order, so the result does not depend on scheduling.  out[F][4] = x, y, v, count.
__global__ void __launch_bounds__(256)
cop_frames_kernel(int64_t n_det, int64_t F, const double* __restrict__ maps, const double* __restrict__ cx,
                  const double* __restrict__ cy, double frac, double* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  __shared__ double red[4][8][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t f = (int64_t)blockIdx.x * 32 + lane;
  const bool ok = f < F;
  double vmax = -INFINITY, vmin = INFINITY;
  if (ok) {
    for (int64_t r = warp; r < n_det; r += 8) {
      const double v = maps[r * F + f];
      vmax = fmax(vmax, v);
      vmin = fmin(vmin, v);
    }
  }
  red[0][warp][lane] = vmax;
  red[1][warp][lane] = vmin;
  __syncthreads();
#pragma unroll
  for (int w = 0; w < 8; ++w) {
    vmax = fmax(vmax, red[0][w][lane]);
    vmin = fmin(vmin, red[1][w][lane]);
  }
  __syncthreads();
  const double th = vmax * frac + vmin * (1.0 - frac);
  double sx = 0.0, sy = 0.0, sv = 0.0, cnt = 0.0;
  if (ok) {
    for (int64_t r = warp; r < n_det; r += 8) {
      const double v = maps[r * F + f];
      if (v > th) {
        sx += cx[r];
        sy += cy[r];
        sv += v;
        cnt += 1.0;
      }
    }
  }
  red[0][warp][lane] = sx;
  red[1][warp][lane] = sy;
  red[2][warp][lane] = sv;
  red[3][warp][lane] = cnt;
  __syncthreads();
  if (warp == 0 && ok) {
    sx = sy = sv = cnt = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      sx += red[0][w][lane];
      sy += red[1][w][lane];
      sv += red[2][w][lane];
      cnt += red[3][w][lane];
    }
    out[4 * f + 0] = sx / cnt;
    out[4 * f + 1] = sy / cnt;
    out[4 * f + 2] = sv / cnt;
    out[4 * f + 3] = cnt;
  }
}

}  // namespace ceit
