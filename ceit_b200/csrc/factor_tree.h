// Multi-level nested-dissection base stage (multifrontal Cholesky, multi-right-hand-side sweeps and selected
// inverse over the elimination forest of plan.h::TreePlan), written once over a backend `BK` that provides the
// batched dense primitives.  The CUDA backend (api.cu) launches the library's kernels on two streams; a host
// backend (oracle/nd_host_check.cpp, TEST INFRASTRUCTURE) runs the same sequence with plain loops so that the
// symbolic tables and the level schedule can be verified against a dense inverse without a GPU.
//
// Replaces the np.linalg.inv(K_ff) of every forward solve of the reference (CEIT/EFEM.py:111-119).
//
// Per front f (own block A_f: s x s, coupling Bt_f: s x b = own rows x boundary columns, boundary block C_f: b x b):
//   bottom-up   A_f = L L^T, Linv;  Y = Linv Bt;  U_f = C_f - Y^T Y  (update matrix, lower triangle valid);
//               extend-add U_f into the parent's (A, Bt, C) through the relative indices `rel` - pushed by the child
//               level (one pass per sibling index) or, for levels of small fronts (`pull`), pulled by the parent:
//               there the whole front step (pull, factorisation, A^-1, Y, U, W) is ONE kernel per level;
//               right-hand sides: t_f = A_f^-1 r_f,  r[bnd(f)] -= Bt_f^T t_f - only for the ACTIVE fronts (K_0U is zero
//               away from the electrodes, so most leaves have r = 0, t = 0; their rows sit after row rowsA);
//   top-down    Sigma_bb(f) gathered from the parent's selected inverse;  V_f = -W_f Sigma_bb  (W = A^-1 Bt);
//               Sigma_ff = A_f^-1 - V_f W_f^T;   x_f = t_f - W_f x[bnd(f)].
// Lanes: 0 = matrix chain (caller's stream), 1 = right-hand-side chain, 2 = W = A^-1 Bt (needed top-down only, so it is
// kept off both bottom-up chains).
#pragma once
#include <cstdint>

#include "plan.h"

namespace ceit {

template <class T>
struct TreeBufs {
  T *A = nullptr, *Lf = nullptr, *Linv = nullptr, *Ainv = nullptr;   // (n s s) storages
  T *Bt = nullptr, *W = nullptr, *V = nullptr;                        // (n s b) storages (V holds Y during bottom-up)
  T* C = nullptr;                                                      // (n b b) storage: U bottom-up, Sigma_bb top-down
  T* SigPP = nullptr;                                                  // (n s s): selected inverse on the own blocks
  T *R = nullptr, *Xp = nullptr;                                       // N0P x nU: swept right-hand sides, t then x
  T* G = nullptr;                                                      // maxG x nU scratch
};

// Domain-decomposed plans (TreePlan::parts > 1): a rank runs the segments of its own sub-trees and of
// the shared top.  Bottom-up is split at t.n_own_segs: [0, n_own_segs) = the rank-owned segments, then ONE exchange
// (update matrices of the sub-tree roots all-gathered into their C slots, right-hand-side rows of the top and the
// partial S_U all-reduced), then [n_own_segs, H) = the top, replicated.  The top-down passes need no exchange: a
// front reads only its ancestors, which are its own or the top's.
inline bool tree_segment_is_mine(const TreePlan& t, const TreeLevel& lv) { return lv.owner < 0 || lv.owner == t.part; }

template <class T, class BK>
void tree_bottom_up(const TreePlan& t, int64_t nU, const TreeBufs<T>& m, BK& k, T one, T zero, T mone,
                    int32_t seg_from = 0, int32_t seg_to = -1) {
  if (seg_to < 0) seg_to = t.H;
  for (int32_t l = seg_from; l < seg_to; ++l) {
    const TreeLevel& lv = t.lv[l];
    const int64_t s = lv.s, b = lv.b, n = lv.n;
    if (n == 0 || s == 0 || !tree_segment_is_mine(t, lv)) continue;
    const int64_t sA = s * s, sB = s * b, sC = b * b, sX = s * nU, sG = b * nU;
    T *A = m.A + lv.offA, *Lf = m.Lf + lv.offA, *Li = m.Linv + lv.offA, *Ai = m.Ainv + lv.offA;
    T *Bt = m.Bt + lv.offB, *W = m.W + lv.offB, *Y = m.V + lv.offB, *C = m.C + lv.offC;
    T *R = m.R + lv.row0 * nU, *X = m.Xp + lv.row0 * nU;      // rows of the ACTIVE fronts (non-zero right-hand side)
    const int na = lv.na;
    // ---- matrix chain ----
    const bool fused = lv.fusable && k.can_fuse();
    if (fused) {
      // ONE kernel per level: pull the children's update matrices, factorise, A^-1, Y, U, W (small fronts)
      k.front_fused(0, t, lv, m.A, m.Bt, m.C, m.Ainv, m.W);
      k.after(0, 1);
      for (int pass = 0; pass < lv.passes; ++pass) k.extend_add(0, t, lv, pass, C, m.A, m.Bt, m.C);
    } else {
      if (lv.pull) k.pull_children(0, t, lv, m.A, m.Bt, m.C);
      k.potrf(0, s, A, Lf, Li, (int)n);
      k.after(0, 1);
      if (b > 0) {
        k.gemm(0, 0, 0, s, b, s, one, Li, s, sA, Bt, b, sB, zero, Y, b, sB, (int)n, 0);        // Y = Linv Bt
        k.gemm(0, 1, 0, b, b, s, mone, Y, b, sB, Y, b, sB, one, C, b, sC, (int)n, 1);         // U = C - Y^T Y (lower)
        for (int pass = 0; pass < lv.passes; ++pass) k.extend_add(0, t, lv, pass, C, m.A, m.Bt, m.C);
      }
      // ---- right-hand-side chain ----
      k.gemm(1, 1, 0, s, s, s, one, Li, s, sA, Li, s, sA, zero, Ai, s, sA, (int)n, 0);         // A^-1 = Linv^T Linv
      if (b > 0) {
        k.after(1, 2);
        k.gemm(2, 0, 0, s, b, s, one, Ai, s, sA, Bt, b, sB, zero, W, b, sB, (int)n, 0);        // W = A^-1 Bt
      }
    }
    if (na > 0) {
      k.gemm(1, 0, 0, s, nU, s, one, Ai, s, sA, R, nU, sX, zero, X, nU, sX, na, 0);            // t = A^-1 r
      if (b > 0) {
        k.gemm(1, 1, 0, b, nU, s, one, Bt, b, sB, X, nU, sX, zero, m.G, nU, sG, na, 0);        // Bt^T t
        k.rhs_scatter(1, t, lv, nU, m.G, m.R);
      }
    }
  }
}

// Requires SigPP == A^-1 (copied by the caller after the bottom-up pass) and lanes 0 and 1 ordered after lane 2 (W).
template <class T, class BK>
void tree_top_down(const TreePlan& t, int64_t nU, const TreeBufs<T>& m, BK& k, T one, T zero, T mone) {
  for (int32_t l = t.H - 1; l >= 0; --l) {
    const TreeLevel& lv = t.lv[l];
    const int64_t s = lv.s, b = lv.b, n = lv.n;
    if (n == 0 || s == 0 || b == 0 || !tree_segment_is_mine(t, lv)) continue;
    const int64_t sA = s * s, sB = s * b, sC = b * b, sX = s * nU, sG = b * nU;
    T *W = m.W + lv.offB, *V = m.V + lv.offB, *C = m.C + lv.offC, *Sg = m.SigPP + lv.offA;
    T *X = m.Xp + lv.row0 * nU, *Xi = m.Xp + lv.row0i * nU;
    const int na = lv.na, ni = (int)n - lv.na;
    // ---- selected inverse ----
    k.extend_gather(0, t, lv, m.SigPP, m.V, m.C);                                               // C_f = Sigma_bb(f)
    k.gemm(0, 0, 0, s, b, b, mone, W, b, sB, C, b, sC, zero, V, b, sB, (int)n, 0);              // V = -W Sigma_bb
    k.gemm(0, 0, 1, s, s, b, mone, V, b, sB, W, b, sB, one, Sg, s, sA, (int)n, 0);              // Sigma_ff = A^-1 - V W^T
    // ---- solution ----
    k.rows_gather(1, t, lv, nU, m.Xp, m.G);
    if (na > 0) k.gemm(1, 0, 0, s, nU, b, mone, W, b, sB, m.G, nU, sG, one, X, nU, sX, na, 0);  // x = t - W x_bnd
    if (ni > 0)                                                                                 // t = 0: x = -W x_bnd
      k.gemm(1, 0, 0, s, nU, b, mone, W + (int64_t)na * sB, b, sB, m.G + (int64_t)na * sG, nU, sG, zero, Xi, nU, sX, ni,
             0);
  }
}

// The solution half of the top-down pass on its own, for ANOTHER block of right-hand sides held in `X` (N0P x nU, rows
// of the active fronts = t, anything elsewhere): x_f = t_f - W_f x[bnd(f)] (t_f = 0 for inactive fronts).
// Used for Y_T = X T without the N x nU x nU product: (t - W x_bnd) T = (t T) - W (x_bnd T), so the same sweep applied
// to t T (a product over the few ACTIVE rows only) yields X T for every row.
template <class T, class BK>
void tree_back_sweep(const TreePlan& t, int64_t nU, const T* Wall, T* G, T* X, BK& k, int lane, T zero, T one, T mone) {
  for (int32_t l = t.H - 1; l >= 0; --l) {
    const TreeLevel& lv = t.lv[l];
    const int64_t s = lv.s, b = lv.b, n = lv.n;
    if (n == 0 || s == 0 || !tree_segment_is_mine(t, lv)) continue;
    const int64_t sB = s * b, sX = s * nU, sG = b * nU;
    const int na = lv.na, ni = (int)n - lv.na;
    T *Xa = X + lv.row0 * nU, *Xi = X + lv.row0i * nU;
    if (b == 0) {                                   // roots: x = t (inactive roots: x = 0)
      if (ni > 0) k.zero_rows(lane, Xi, (int64_t)ni * s * nU);
      continue;
    }
    const T* W = Wall + lv.offB;
    k.rows_gather(lane, t, lv, nU, X, G);
    if (na > 0) k.gemm(lane, 0, 0, s, nU, b, mone, W, b, sB, G, nU, sG, one, Xa, nU, sX, na, 0);
    if (ni > 0)
      k.gemm(lane, 0, 0, s, nU, b, mone, W + (int64_t)na * sB, b, sB, G + (int64_t)na * sG, nU, sG, zero, Xi, nU, sX, ni, 0);
  }
}

}  // namespace ceit
