// TEST INFRASTRUCTURE (not part of the product; only tests/ build and call this).
// Host backend for ceit_b200/csrc/factor_tree.h: runs the multi-level nested-dissection base stage of the
// library - the SAME symbolic tables (plan.cpp) and the SAME level schedule (factor_tree.h) as the CUDA path - with
// plain loops, so that X = K00^-1 K0U, S_U and the selected inverse on the mesh pattern can be compared with a dense
// numpy inverse on the CPU.  The reference has no counterpart (it inverts the dense matrix, CEIT/EFEM.py:111-119).
#include <cmath>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../ceit_b200/csrc/factor_tree.h"
#include "../ceit_b200/csrc/plan.h"

using namespace ceit;

namespace {
struct HostBK {
  int bad_pivot = 0;
  void after(int, int) {}
  void gemm(int, int opA, int opB, int64_t M, int64_t N, int64_t K, double alpha, const double* A, int64_t lda,
            int64_t sA, const double* B, int64_t ldb, int64_t sB, double beta, double* C, int64_t ldc, int64_t sC,
            int batch, int lower_only) {
    for (int z = 0; z < batch; ++z) {
      const double *a = A + z * sA, *b = B + z * sB;
      double* c = C + z * sC;
      for (int64_t i = 0; i < M; ++i)
        for (int64_t j = 0; j < N; ++j) {
          if (lower_only && j > i) continue;      // entries above the diagonal are NOT produced (the GPU skips whole tiles)
          double acc = 0.0;
          for (int64_t q = 0; q < K; ++q)
            acc += (opA ? a[q * lda + i] : a[i * lda + q]) * (opB ? b[j * ldb + q] : b[q * ldb + j]);
          c[i * ldc + j] = alpha * acc + (beta == 0.0 ? 0.0 : beta * c[i * ldc + j]);
        }
    }
  }
  void potrf(int, int64_t n, double* A, double* L, double* Li, int batch) {
    for (int z = 0; z < batch; ++z) {
      double *a = A + z * n * n, *l = L + z * n * n, *x = Li + z * n * n;
      for (int64_t j = 0; j < n; ++j) {                     // lower triangle of A only
        double d = a[j * n + j];
        for (int64_t q = 0; q < j; ++q) d -= l[j * n + q] * l[j * n + q];
        if (!(d > 0.0)) bad_pivot = 1;
        l[j * n + j] = std::sqrt(d);
        for (int64_t i = j + 1; i < n; ++i) {
          double v = a[i * n + j];
          for (int64_t q = 0; q < j; ++q) v -= l[i * n + q] * l[j * n + q];
          l[i * n + j] = v / l[j * n + j];
        }
      }
      for (int64_t c = 0; c < n; ++c)                       // X = L^-1 by forward substitution
        for (int64_t i = c; i < n; ++i) {
          double v = (i == c) ? 1.0 : 0.0;
          for (int64_t q = c; q < i; ++q) v -= l[i * n + q] * x[q * n + c];
          x[i * n + c] = v / l[i * n + i];
        }
    }
  }
  bool fuse = true;
  bool can_fuse() const { return fuse; }
  // destination-centric pull of the children's update matrices into (A, Bt, C) of every front of the level (the
  // loop structure of pull_children_kernel: one "thread" per entry of the parent's front index space)
  void pull_children(int, const TreePlan& t, const TreeLevel& lv, double* A, double* Bt, double* C) {
    const int64_t s = lv.s, b = lv.b, span = s + b;
    for (int32_t f = 0; f < lv.n; ++f) {
      const int32_t g = lv.f0 + f;
      double *a_ = A + lv.offA + (int64_t)f * s * s, *bt = Bt + lv.offB + (int64_t)f * s * b;
      double* c_ = C + lv.offC + (int64_t)f * b * b;
      for (int64_t ra = 0; ra < span; ++ra)
        for (int64_t rc = 0; rc < span; ++rc) {
          if (ra >= s && rc < s) continue;
          double acc = 0.0;
          bool hit = false;
          for (int32_t q = t.child_ptr[g]; q < t.child_ptr[g + 1]; ++q) {
            const int32_t ch = t.child_ids[q];
            const int32_t* inv = t.inv_rel.data() + t.f_invOff[ch];
            const int64_t ia = inv[ra], ic = inv[rc];
            if (ia < 0 || ic < 0) continue;
            const int64_t cb = t.f_b[ch];
            const double* u = C + t.f_offC[ch];
            acc += (ia >= ic) ? u[ia * cb + ic] : u[ic * cb + ia];
            hit = true;
          }
          if (!hit) continue;
          if (ra < s) {
            if (rc < s) a_[ra * s + rc] += acc;
            else bt[ra * b + (rc - s)] += acc;
          } else {
            c_[(ra - s) * b + (rc - s)] += acc;
          }
        }
    }
  }
  // the fused front step of a pull level, with the same outputs as the CUDA kernel: A^-1, W, U (in C, lower) and
  // the updated Bt; L / Linv are not kept
  void front_fused(int, const TreePlan& t, const TreeLevel& lv, double* A, double* Bt, double* C, double* Ainv,
                   double* W) {
    const int64_t s = lv.s, b = lv.b, n = lv.n;
    pull_children(0, t, lv, A, Bt, C);
    std::vector<double> L(n * s * s, 0.0), X(n * s * s, 0.0), Y(std::max<int64_t>(n * s * b, 1), 0.0);
    potrf(0, s, A + lv.offA, L.data(), X.data(), (int)n);
    gemm(0, 1, 0, s, s, s, 1.0, X.data(), s, s * s, X.data(), s, s * s, 0.0, Ainv + lv.offA, s, s * s, (int)n, 0);
    if (b > 0) {
      gemm(0, 0, 0, s, b, s, 1.0, X.data(), s, s * s, Bt + lv.offB, b, s * b, 0.0, Y.data(), b, s * b, (int)n, 0);
      gemm(0, 1, 0, b, b, s, -1.0, Y.data(), b, s * b, Y.data(), b, s * b, 1.0, C + lv.offC, b, b * b, (int)n, 1);
      gemm(0, 0, 0, s, b, s, 1.0, Ainv + lv.offA, s, s * s, Bt + lv.offB, b, s * b, 0.0, W + lv.offB, b, s * b, (int)n, 0);
    }
  }
  void extend_add(int, const TreePlan& t, const TreeLevel& lv, int pass, const double* U, double* A, double* Bt,
                  double* C) {
    const int64_t b = lv.b;
    for (int32_t f = 0; f < lv.n; ++f) {
      const int32_t g = lv.f0 + f;
      if (t.sib[g] != pass || t.parent[g] < 0 || !t.push[g]) continue;
      const int32_t* rel = t.rel.data() + lv.offR + (int64_t)f * b;
      const double* u = U + (int64_t)f * b * b;
      const int64_t ps = t.ps[g], pb = t.pb[g];
      for (int64_t a = 0; a < b; ++a)
        for (int64_t c = 0; c < b; ++c) {
          const int64_t ra = rel[a], rc = rel[c];
          if (ra < 0 || rc < 0) continue;
          const double v = (a >= c) ? u[a * b + c] : u[c * b + a];
          if (ra < ps) {
            if (rc < ps) A[t.pA[g] + ra * ps + rc] += v;
            else Bt[t.pB[g] + ra * pb + (rc - ps)] += v;
          } else if (rc >= ps) {
            C[t.pC[g] + (ra - ps) * pb + (rc - ps)] += v;
          }
        }
    }
  }
  void extend_gather(int, const TreePlan& t, const TreeLevel& lv, const double* Sg, const double* V, double* C) {
    const int64_t b = lv.b;
    for (int32_t f = 0; f < lv.n; ++f) {
      const int32_t g = lv.f0 + f;
      const int32_t* rel = t.rel.data() + lv.offR + (int64_t)f * b;
      double* out = C + lv.offC + (int64_t)f * b * b;
      const int64_t ps = t.ps[g], pb = t.pb[g];
      for (int64_t a = 0; a < b; ++a)
        for (int64_t c = 0; c < b; ++c) {
          const int64_t ra = rel[a], rc = rel[c];
          double v = 0.0;
          if (ra >= 0 && rc >= 0 && t.parent[g] >= 0) {
            if (ra < ps && rc < ps) v = Sg[t.pA[g] + ra * ps + rc];
            else if (ra < ps) v = V[t.pB[g] + ra * pb + (rc - ps)];
            else if (rc < ps) v = V[t.pB[g] + rc * pb + (ra - ps)];
            else v = C[t.pC[g] + (ra - ps) * pb + (rc - ps)];
          }
          out[a * b + c] = v;
        }
    }
  }
  void rhs_scatter(int, const TreePlan& t, const TreeLevel& lv, int64_t nU, const double* G, double* R) {
    for (int64_t q = lv.rc0; q < lv.rc0 + lv.rc_n; ++q)
      for (int64_t c = 0; c < nU; ++c) {
        double acc = 0.0;
        for (int32_t s = t.rc_ptr[q]; s < t.rc_ptr[q + 1]; ++s) acc += G[(int64_t)t.rc_src[s] * nU + c];
        R[(int64_t)t.rc_dst[q] * nU + c] -= acc;
      }
  }
  void zero_rows(int, double* ptr, int64_t count) { std::memset(ptr, 0, sizeof(double) * (size_t)count); }
  void rows_gather(int, const TreePlan& t, const TreeLevel& lv, int64_t nU, const double* Xp, double* G) {
    for (int64_t r = 0; r < (int64_t)lv.n * lv.b; ++r) {
      const int32_t src = t.bnd_row[lv.offR + r];
      for (int64_t c = 0; c < nU; ++c) G[r * nU + c] = src < 0 ? 0.0 : Xp[(int64_t)src * nU + c];
    }
  }
};
thread_local std::string g_msg;
}  // namespace

extern "C" {

const char* nd_host_last_error() { return g_msg.c_str(); }

// K: dense symmetric N x N (row-major, the reference's doubled matrix or any SPD-on-F0 matrix with the mesh pattern).
// Outputs (caller-allocated): info i64[8] = {nU, N0, H, nF, N0P, max s, max b, bad_pivot}; pos i64[N] (compact
// positions), U i64[nU]; X f64[N0 * nU] (rows in compact position order) = K00^-1 K0U; SU f64[nU*nU];
// g0 f64[6 E] selected-inverse entries of the element node pairs (0,0),(1,1),(2,2),(0,1),(1,2),(2,0) (0 if a node is
// an electrode node).  Any output pointer may be NULL (info first to size the others).
int nd_host_check(const double* nodes, const int64_t* elem, int64_t N, int64_t E, const int64_t* el_ptr,
                  const int64_t* el_elems, int64_t L, int64_t leaf_cap, const double* K, int64_t* info, int64_t* pos,
                  int64_t* U, double* X, double* SU, double* g0, int mode, int ways) {
  // mode: 0 = pull plan, fusable levels run the fused front step, 1 = pull plan, separate steps everywhere (default of
  // the CUDA backend), 2 = push plan (one extend-add pass per sibling index); ways = 2 / 4 cuts per front
  try {
    HostPlan h;
    build_host_plan(h, nodes, elem, N, E, el_ptr, el_elems, L, 0, leaf_cap, /*nd_tree=*/true, /*nd_pull=*/mode != 2 ? 1 : 0, ways);
    if (!h.use_tree) throw std::runtime_error("mesh too small for the tree plan");
    const TreePlan& t = h.tree;
    int64_t smax = 0, bmax = 0;
    for (auto& lv : t.lv) { smax = std::max<int64_t>(smax, lv.s); bmax = std::max<int64_t>(bmax, lv.b); }
    if (info) {
      info[0] = h.nU; info[1] = h.N0; info[2] = t.H; info[3] = t.nF; info[4] = h.N0P; info[5] = smax; info[6] = bmax;
      info[7] = 0;
      int64_t npull = 0;
      for (auto& lv : t.lv) npull += lv.pull;
      info[6] = bmax + 100000 * npull;           // packed: number of pull levels in the upper digits
    }
    if (pos) for (int64_t v = 0; v < N; ++v) pos[v] = h.pos[v];
    if (U) for (int64_t u = 0; u < h.nU; ++u) U[u] = h.U[u];
    if (!K || !X || !SU || !g0) return 0;
    const int64_t nU = h.nU, N0P = h.N0P;
    std::vector<double> A(t.szA, 0.0), Lf(t.szA, 0.0), Li(t.szA, 0.0), Ai(t.szA, 0.0), Bt(t.szB, 0.0), W(t.szB, 0.0);
    std::vector<double> Sel(t.szA + t.szB, 0.0), C(t.szC, 0.0), R(N0P * nU, 0.0), Xp(N0P * nU, 0.0);
    for (int64_t q = t.rowsA * nU; q < N0P * nU; ++q) R[q] = Xp[q] = std::nan("");   // inactive rows: never read before written
    std::vector<double> G(std::max<int64_t>(t.maxG, 1) * nU, 0.0), KUU(nU * nU, 0.0);
    // identity on the padding slots (pad_identity_kernel) and the CSR scatter (scatter_dense_kernel)
    for (int64_t d : t.pad_diag) A[d] = 1.0;
    for (int64_t r = 0; r < N; ++r)
      for (int32_t s = h.rowptr[r]; s < h.rowptr[r + 1]; ++s) {
        const double v = K[r * N + h.colidx[s]];
        const int64_t d = h.slot_dst[s];
        switch (h.slot_kind[s]) {
          case 5: A[d] = v; break;
          case 6: Bt[d] = v; break;
          case 3: R[d] = v; break;
          case 4: KUU[d] = v; break;
          default: break;
        }
      }
    TreeBufs<double> m;
    m.A = A.data(); m.Lf = Lf.data(); m.Linv = Li.data(); m.Ainv = Ai.data(); m.Bt = Bt.data(); m.W = W.data();
    m.SigPP = Sel.data(); m.V = Sel.data() + t.szA; m.C = C.data(); m.R = R.data(); m.Xp = Xp.data(); m.G = G.data();
    HostBK bk;
    bk.fuse = (mode == 0);
    tree_bottom_up<double>(t, nU, m, bk, 1.0, 0.0, -1.0);
    // S_U = K_UU - r^T t (the swept right-hand sides against t = A^-1 r, all levels at once)
    for (int64_t a = 0; a < nU; ++a)
      for (int64_t b = 0; b < nU; ++b) {
        double acc = 0.0;
        for (int64_t r = 0; r < t.rowsA; ++r) acc += R[r * nU + a] * Xp[r * nU + b];
        SU[a * nU + b] = KUU[a * nU + b] - acc;
      }
    std::memcpy(Sel.data(), Ai.data(), sizeof(double) * t.szA);
    std::vector<double> tAct(Xp.begin(), Xp.begin() + t.rowsA * nU);       // t of the active rows (the sweep is in place)
    tree_top_down<double>(t, nU, m, bk, 1.0, 0.0, -1.0);
    // second backward sweep on t T' (T' = a fixed dense nU x nU matrix): must equal X T' on every row
    {
      std::vector<double> Tp(nU * nU), YP(N0P * nU, std::nan(""));
      for (int64_t a = 0; a < nU; ++a)
        for (int64_t b = 0; b < nU; ++b) Tp[a * nU + b] = 1.0 / (1.0 + (double)((a > b) ? a - b : b - a)) + (a == b ? 2.0 : 0.0);
      bk.gemm(0, 0, 0, t.rowsA, nU, nU, 1.0, tAct.data(), nU, 0, Tp.data(), nU, 0, 0.0, YP.data(), nU, 0, 1, 0);
      tree_back_sweep<double>(t, nU, W.data(), G.data(), YP.data(), bk, 1, 0.0, 1.0, -1.0);
      double worst = 0.0, scale = 0.0;
      for (int64_t q = 0; q < h.N0; ++q) {
        const int64_t r = h.row_map[q];
        for (int64_t b = 0; b < nU; ++b) {
          double ref = 0.0;
          for (int64_t a = 0; a < nU; ++a) ref += Xp[r * nU + a] * Tp[a * nU + b];
          worst = std::max(worst, std::fabs(ref - YP[r * nU + b]));
          scale = std::max(scale, std::fabs(ref));
        }
      }
      if (!(worst <= 1e-11 * std::max(scale, 1.0))) throw std::runtime_error("second backward sweep differs from X T'");
    }
    for (int64_t q = 0; q < h.N0; ++q)
      std::memcpy(X + q * nU, Xp.data() + (int64_t)h.row_map[q] * nU, sizeof(double) * nU);
    for (int64_t q = 0; q < 6 * E; ++q) g0[q] = h.g0_off[q] < 0 ? 0.0 : Sel[h.g0_off[q]];
    if (info) info[7] = bk.bad_pivot;
    return 0;
  } catch (const std::exception& e) {
    g_msg = e.what();
    return -1;
  }
}

// ---- domain-decomposed base stage (TreePlan::parts > 1): the ranks are simulated one after the other on the host,
// the NCCL exchange of the product (all-gather of the sub-tree roots' update matrices, all-reduce of the top's
// right-hand-side rows and of the partial S_U) is done with plain copies / sums.  Outputs per rank r:
//   n0c i64[parts], pos i64[parts * N] (compact positions, -1 = foreign), X f64[parts * N0 * nU] (first n0c[r] rows of
//   slab r), SU f64[parts * nU * nU], g0 f64[6 E] (from each element's owner), owner i64[E] (rank of each element; -2
//   if owned by none or by several = an error the test reports).
namespace {
struct RankState {
  HostPlan h;
  std::vector<double> A, Lf, Li, Ai, Bt, W, Sel, C, R, Xp, G, KUU, P;
  TreeBufs<double> m;
  HostBK bk;
};
}  // namespace

int nd_host_check_parts(const double* nodes, const int64_t* elem, int64_t N, int64_t E, const int64_t* el_ptr,
                        const int64_t* el_elems, int64_t L, int64_t leaf_cap, const double* K, int parts, int ways,
                        int64_t* info, int64_t* n0c, int64_t* pos, double* X, double* SU, double* g0, int64_t* owner) {
  try {
    std::vector<RankState> rk(parts);
    for (int r = 0; r < parts; ++r) {
      RankState& s = rk[r];
      build_host_plan(s.h, nodes, elem, N, E, el_ptr, el_elems, L, 0, leaf_cap, true, 1, ways, parts, r);
      if (!s.h.use_tree) throw std::runtime_error("mesh too small for the tree plan");
      build_patches(s.h, 24, 32);
    }
    const TreePlan& t0 = rk[0].h.tree;
    const int64_t nU = rk[0].h.nU, N0P = rk[0].h.N0P, N0 = rk[0].h.N0;
    if (info) {
      info[0] = nU; info[1] = N0; info[2] = t0.H; info[3] = t0.nF; info[4] = N0P; info[5] = t0.n_own_segs;
      info[6] = t0.roots_per_rank; info[7] = 0; info[8] = t0.dcut; info[9] = t0.rowsA - t0.ra_top0;
    }
    for (int r = 0; r < parts; ++r) {
      if (n0c) n0c[r] = rk[r].h.N0c;
      if (pos) for (int64_t v = 0; v < N; ++v) pos[r * N + v] = rk[r].h.pos[v];
    }
    if (owner) {
      for (int64_t e = 0; e < E; ++e) owner[e] = -2;
      for (int r = 0; r < parts; ++r) {
        const HostPlan& h = rk[r].h;
        for (int64_t e = 0; e < E; ++e)
          if (h.elem_own[e]) owner[e] = (owner[e] == -2) ? r : -3;
        // every owned element sits in exactly one patch of its rank, foreign elements in none
        std::vector<int> seen(E, 0);
        for (int32_t e : h.patch_elems) seen[e]++;
        for (int64_t e = 0; e < E; ++e)
          if (seen[e] != (h.elem_own[e] ? 1 : 0)) throw std::runtime_error("patches do not cover exactly the owned elements");
        for (int32_t q : h.patch_nodes)
          if (q < 0 || q >= h.Nc) throw std::runtime_error("a patch node has no row on its rank");
      }
    }
    if (!K || !X || !SU || !g0) return 0;
    // phase A on every rank: scatter, own segments bottom-up, partial S_U over the rank's active rows
    for (int r = 0; r < parts; ++r) {
      RankState& s = rk[r];
      const HostPlan& h = s.h;
      const TreePlan& t = h.tree;
      s.A.assign(t.szA, 0.0); s.Lf.assign(t.szA, 0.0); s.Li.assign(t.szA, 0.0); s.Ai.assign(t.szA, 0.0);
      s.Bt.assign(t.szB, 0.0); s.W.assign(t.szB, 0.0); s.Sel.assign(t.szA + t.szB, 0.0); s.C.assign(t.szC, 0.0);
      s.R.assign(N0P * nU, 0.0); s.Xp.assign(N0P * nU, 0.0);
      s.G.assign(std::max<int64_t>(t.maxG, 1) * nU, 0.0); s.KUU.assign(nU * nU, 0.0); s.P.assign(nU * nU, 0.0);
      for (int64_t d : t.pad_diag) s.A[d] = 1.0;
      for (int64_t q = 0; q < N; ++q)
        for (int32_t c = h.rowptr[q]; c < h.rowptr[q + 1]; ++c) {
          const double v = K[q * N + h.colidx[c]];
          const int64_t d = h.slot_dst[c];
          switch (h.slot_kind[c]) {
            case 5: s.A[d] = v; break;
            case 6: s.Bt[d] = v; break;
            case 3: s.R[d] = v; break;
            case 4: s.KUU[d] = v; break;
            default: break;
          }
        }
      if (r != 0)                                            // the assembled rows of the top count once in the sum
        std::fill(s.R.begin() + t.ra_top0 * nU, s.R.begin() + t.rowsA * nU, 0.0);
      // foreign rows must never be read: poison them
      for (int q = 0; q < parts; ++q)
        if (q != r)
          for (int64_t x = t.ra0[q] * nU; x < t.ra1[q] * nU; ++x) s.R[x] = s.Xp[x] = std::nan("");
      s.m.A = s.A.data(); s.m.Lf = s.Lf.data(); s.m.Linv = s.Li.data(); s.m.Ainv = s.Ai.data(); s.m.Bt = s.Bt.data();
      s.m.W = s.W.data(); s.m.SigPP = s.Sel.data(); s.m.V = s.Sel.data() + t.szA; s.m.C = s.C.data();
      s.m.R = s.R.data(); s.m.Xp = s.Xp.data(); s.m.G = s.G.data();
      s.bk.fuse = false;
      tree_bottom_up<double>(t, nU, s.m, s.bk, 1.0, 0.0, -1.0, 0, t.n_own_segs);
      for (int64_t a = 0; a < nU; ++a)
        for (int64_t b = 0; b < nU; ++b) {
          double acc = 0.0;
          for (int64_t q = t.ra0[r]; q < t.ra1[r]; ++q) acc += s.R[q * nU + a] * s.Xp[q * nU + b];
          s.P[a * nU + b] = acc;
        }
    }
    // the exchange
    for (int r = 0; r < parts; ++r)
      for (int32_t k = 0; k < t0.roots_per_rank; ++k) {
        const int64_t b = t0.exp_b[(size_t)r * t0.roots_per_rank + k], off = t0.exp_offC[(size_t)r * t0.roots_per_rank + k];
        for (int q = 0; q < parts; ++q)
          if (q != r) std::memcpy(rk[q].C.data() + off, rk[r].C.data() + off, sizeof(double) * (size_t)(b * b));
      }
    {
      std::vector<double> sumR((t0.rowsA - t0.ra_top0) * nU, 0.0), sumP(nU * nU, 0.0);
      for (int r = 0; r < parts; ++r) {
        for (size_t x = 0; x < sumR.size(); ++x) sumR[x] += rk[r].R[t0.ra_top0 * nU + x];
        for (size_t x = 0; x < sumP.size(); ++x) sumP[x] += rk[r].P[x];
      }
      for (int r = 0; r < parts; ++r) {
        std::copy(sumR.begin(), sumR.end(), rk[r].R.begin() + t0.ra_top0 * nU);
        rk[r].P = sumP;
      }
    }
    // phase B on every rank
    for (int r = 0; r < parts; ++r) {
      RankState& s = rk[r];
      const HostPlan& h = s.h;
      const TreePlan& t = h.tree;
      tree_bottom_up<double>(t, nU, s.m, s.bk, 1.0, 0.0, -1.0, t.n_own_segs, t.H);
      double* su = SU + (int64_t)r * nU * nU;
      for (int64_t a = 0; a < nU; ++a)
        for (int64_t b = 0; b < nU; ++b) {
          double acc = 0.0;
          for (int64_t q = t.ra_top0; q < t.rowsA; ++q) acc += s.R[q * nU + a] * s.Xp[q * nU + b];
          su[a * nU + b] = s.KUU[a * nU + b] - s.P[a * nU + b] - acc;
        }
      std::memcpy(s.Sel.data(), s.Ai.data(), sizeof(double) * t.szA);
      tree_top_down<double>(t, nU, s.m, s.bk, 1.0, 0.0, -1.0);
      for (int64_t q = 0; q < h.N0c; ++q)
        std::memcpy(X + ((int64_t)r * N0 + q) * nU, s.Xp.data() + (int64_t)h.row_map[q] * nU, sizeof(double) * nU);
      for (int64_t e = 0; e < E; ++e)
        if (h.elem_own[e])
          for (int q = 0; q < 6; ++q) g0[6 * e + q] = h.g0_off[6 * e + q] < 0 ? 0.0 : s.Sel[h.g0_off[6 * e + q]];
      if (info) info[7] |= s.bk.bad_pivot;
    }
    return 0;
  } catch (const std::exception& e) {
    g_msg = e.what();
    return -1;
  }
}
}
