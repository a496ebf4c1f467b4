// Host-side symbolic analysis for the CEIT hot path (pure C++).
//
// What the reference recomputes inside every one of its L*E forward solves - the Dirichlet node
// set of the excitation electrode (CEIT/EFEM.py:79-86), the row/column swaps that move it to the
// tail (EFEM.py:87-96,136-142), the node->element->electrode averaging (CEIT/FEMBasic.py:118-137)
// - is turned here, once per mesh, into static index tables.  Nothing below is a translation of
// reference code: the reference has no ordering, no sparse pattern and no electrode-node union.
#include "plan.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <stdexcept>

namespace ceit {

namespace {

// BFS over nodes with keep[v] && committed[v] < 0; fills `levels` (each sorted ascending).
void bfs_levels(const std::vector<int32_t>& rowptr, const std::vector<int32_t>& colidx,
                const std::vector<uint8_t>& keep, const std::vector<int32_t>& committed,
                std::vector<int32_t>& stamp, int32_t stamp_val, int32_t start,
                std::vector<std::vector<int32_t>>& levels) {
  levels.clear();
  std::vector<int32_t> frontier{start};
  stamp[start] = stamp_val;
  while (!frontier.empty()) {
    levels.push_back(frontier);
    std::vector<int32_t> next;
    for (int32_t v : frontier) {
      for (int32_t q = rowptr[v]; q < rowptr[v + 1]; ++q) {
        int32_t w = colidx[q];
        if (keep[w] && committed[w] < 0 && stamp[w] != stamp_val) {
          stamp[w] = stamp_val;
          next.push_back(w);
        }
      }
    }
    std::sort(next.begin(), next.end());
    frontier.swap(next);
  }
}

void build_tree_plan(HostPlan& p, int64_t leaf_cap, int pull_mode, int ways, int parts, int part);

}  // namespace

void build_host_plan(HostPlan& p, const double* nodes, const int64_t* elem, int64_t N, int64_t E,
                     const int64_t* el_ptr, const int64_t* el_elems, int64_t L,
                     int64_t target_block, int64_t nd_interior, bool nd_tree, int nd_pull, int nd_ways, int parts,
                     int part) {
  if (N <= 0 || E <= 0 || L <= 0) throw std::runtime_error("plan: N, E, L must be positive");
  if (parts < 1 || part < 0 || part >= parts) throw std::runtime_error("plan: invalid domain decomposition (parts, part)");
  if (N > (int64_t)2000000000 / 16 || E > (int64_t)2000000000 / 16)
    throw std::runtime_error("plan: mesh too large for 32-bit indices");
  if (target_block <= 0) target_block = 192;
  p.N = N;
  p.E = E;
  p.L = L;
  p.nodes.assign(nodes, nodes + 2 * N);
  p.elem.resize(3 * E);
  for (int64_t k = 0; k < 3 * E; ++k) {
    if (elem[k] < 0 || elem[k] >= N) throw std::runtime_error("plan: element node index out of range");
    p.elem[k] = (int32_t)elem[k];
  }
  // ---- electrodes: element lists, node sets, union U, weights --------------------------------
  p.el_ptr.resize(L + 1);
  for (int64_t i = 0; i <= L; ++i) p.el_ptr[i] = (int32_t)el_ptr[i];
  if (el_ptr[0] != 0) throw std::runtime_error("plan: el_ptr[0] must be 0");
  p.el_elems.resize(el_ptr[L]);
  for (int64_t k = 0; k < el_ptr[L]; ++k) {
    if (el_elems[k] < 0 || el_elems[k] >= E) throw std::runtime_error("plan: electrode element index out of range");
    p.el_elems[k] = (int32_t)el_elems[k];
  }
  p.u_of.assign(N, -1);
  p.U.clear();
  std::vector<std::vector<int32_t>> node_sets(L);
  for (int64_t i = 0; i < L; ++i) {
    if (el_ptr[i + 1] <= el_ptr[i]) throw std::runtime_error("No element is selected, please check the input");
    auto& s = node_sets[i];
    for (int32_t q = p.el_ptr[i]; q < p.el_ptr[i + 1]; ++q)
      for (int a = 0; a < 3; ++a) s.push_back(p.elem[3 * p.el_elems[q] + a]);
    std::sort(s.begin(), s.end());
    s.erase(std::unique(s.begin(), s.end()), s.end());
    for (int32_t n : s)
      if (p.u_of[n] < 0) {
        p.u_of[n] = (int32_t)p.U.size();
        p.U.push_back(n);
      }
  }
  p.nU = (int64_t)p.U.size();
  p.N0 = N - p.nU;
  if (p.N0 <= 0) throw std::runtime_error("plan: every node belongs to an electrode");
  p.N0c = p.N0;
  p.Nc = N;
  p.elem_own.assign(E, 1);
  p.inB.assign(L * p.nU, 0);
  for (int64_t i = 0; i < L; ++i)
    for (int32_t n : node_sets[i]) p.inB[i * p.nU + p.u_of[n]] = 1;
  // weights: electrode_pot[m] = mean over its elements of mean over the 3 node amplitudes
  p.we_ptr.assign(L + 1, 0);
  p.we_u.clear();
  p.we_w.clear();
  {
    std::vector<double> acc(p.nU, 0.0);
    for (int64_t m = 0; m < L; ++m) {
      int32_t cnt = p.el_ptr[m + 1] - p.el_ptr[m];
      for (int32_t n : node_sets[m]) acc[p.u_of[n]] = 0.0;
      for (int32_t q = p.el_ptr[m]; q < p.el_ptr[m + 1]; ++q)
        for (int a = 0; a < 3; ++a) acc[p.u_of[p.elem[3 * p.el_elems[q] + a]]] += 1.0;
      for (int32_t n : node_sets[m]) {
        p.we_u.push_back(p.u_of[n]);
        p.we_w.push_back(acc[p.u_of[n]] / (3.0 * cnt));
      }
      p.we_ptr[m + 1] = (int32_t)p.we_u.size();
    }
  }
  // contiguous form (valid when no node is shared between electrodes)
  p.u_contiguous = ((int64_t)p.we_u.size() == p.nU);
  p.ue_ptr.assign(L + 1, 0);
  p.wk.assign(p.nU, 0.0);
  if (p.u_contiguous) {
    for (int64_t m = 0; m < L && p.u_contiguous; ++m) {
      p.ue_ptr[m + 1] = p.we_ptr[m + 1];
      for (int32_t q = p.we_ptr[m]; q < p.we_ptr[m + 1]; ++q) {
        if (p.we_u[q] != q) p.u_contiguous = false;
        else p.wk[q] = p.we_w[q];
      }
    }
  }
  {
    int32_t mx = 0;
    for (int64_t m = 0; m < L; ++m) mx = std::max<int32_t>(mx, (int32_t)node_sets[m].size());
    p.bp = (mx + 7) / 8 * 8;
  }
  // ---- CSR pattern of K: node adjacency incl. diagonal ---------------------------------------
  {
    std::vector<int32_t> deg(N + 1, 0);
    for (int64_t e = 0; e < E; ++e)
      for (int a = 0; a < 3; ++a) deg[p.elem[3 * e + a] + 1] += 3;
    std::vector<int64_t> start(N + 1, 0);
    for (int64_t v = 0; v < N; ++v) start[v + 1] = start[v] + deg[v + 1];
    std::vector<int32_t> tmp(start[N]);
    std::vector<int64_t> fill(start.begin(), start.end() - 1);
    for (int64_t e = 0; e < E; ++e)
      for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) tmp[fill[p.elem[3 * e + a]]++] = p.elem[3 * e + b];
    p.rowptr.assign(N + 1, 0);
    p.colidx.clear();
    p.colidx.reserve(7 * N);
    for (int64_t v = 0; v < N; ++v) {
      auto b = tmp.begin() + start[v], en = tmp.begin() + start[v + 1];
      std::sort(b, en);
      en = std::unique(b, en);
      p.colidx.insert(p.colidx.end(), b, en);
      p.rowptr[v + 1] = (int32_t)p.colidx.size();
    }
    p.nnz = (int64_t)p.colidx.size();
  }
  // ---- gather lists for the deterministic assembly + dense destinations ----------------------
  p.src_ptr.assign(p.nnz + 1, 0);
  p.src.assign(9 * E, 0);
  {
    std::vector<int32_t> slot_of(9 * E);
    for (int64_t e = 0; e < E; ++e)
      for (int a = 0; a < 3; ++a) {
        int32_t r = p.elem[3 * e + a];
        for (int b = 0; b < 3; ++b) {
          int32_t c2 = p.elem[3 * e + b];
          auto it = std::lower_bound(p.colidx.begin() + p.rowptr[r], p.colidx.begin() + p.rowptr[r + 1], c2);
          int32_t s = (int32_t)(it - p.colidx.begin());
          slot_of[9 * e + 3 * a + b] = s;
          p.src_ptr[s + 1]++;
        }
      }
    for (int64_t s = 0; s < p.nnz; ++s) p.src_ptr[s + 1] += p.src_ptr[s];
    std::vector<int32_t> fill(p.src_ptr.begin(), p.src_ptr.end() - 1);
    for (int64_t q = 0; q < 9 * E; ++q) p.src[fill[slot_of[q]]++] = (int32_t)q;  // ascending element order
  }
  if (nd_tree && nd_interior > 0 && p.N0 > std::max<int64_t>(nd_interior, 8)) {
    build_tree_plan(p, std::min<int64_t>(nd_interior, 64), nd_pull, nd_ways, parts, part);
    return;
  }
  if (parts > 1) throw std::runtime_error("plan: the domain decomposition needs the multi-level nested-dissection ordering");
  // ---- level 1 (nested dissection): geometric boxes -> interiors I_p (<= 64 nodes, mutually
  // disconnected) and the separator set C.  A node goes to C when it has an F0 neighbour in a box with
  // a larger id, so every edge between two boxes has its lower endpoint in C. ----------------------
  std::vector<int32_t> interior_of(N, -1);      // node -> interior id, -1 for C / U
  std::vector<std::vector<int32_t>> interiors;
  const int64_t kMaxInterior = 64;
  if (nd_interior > 0 && p.N0 > 4 * kMaxInterior) {
    double xmin = 1e300, xmax = -1e300, ymin = 1e300, ymax = -1e300;
    for (int64_t v = 0; v < N; ++v)
      if (p.u_of[v] < 0) {
        xmin = std::min(xmin, nodes[2 * v]); xmax = std::max(xmax, nodes[2 * v]);
        ymin = std::min(ymin, nodes[2 * v + 1]); ymax = std::max(ymax, nodes[2 * v + 1]);
      }
    int64_t g = (int64_t)std::ceil(std::sqrt((double)p.N0 / (double)nd_interior));
    if (g < 2) g = 2;
    for (int attempt = 0; attempt < 12; ++attempt, g = g + 1 + g / 8) {
      const double wx = (xmax - xmin) / g * (1.0 + 1e-12) + 1e-300, wy = (ymax - ymin) / g * (1.0 + 1e-12) + 1e-300;
      std::vector<int64_t> box(N, -1);
      for (int64_t v = 0; v < N; ++v)
        if (p.u_of[v] < 0) {
          int64_t bx = std::min<int64_t>(g - 1, (int64_t)((nodes[2 * v] - xmin) / wx));
          int64_t by = std::min<int64_t>(g - 1, (int64_t)((nodes[2 * v + 1] - ymin) / wy));
          box[v] = by * g + bx;
        }
      std::vector<uint8_t> is_sep(N, 0);
      for (int64_t v = 0; v < N; ++v) {
        if (box[v] < 0) continue;
        for (int32_t q = p.rowptr[v]; q < p.rowptr[v + 1]; ++q) {
          const int32_t w = p.colidx[q];
          if (box[w] > box[v]) { is_sep[v] = 1; break; }
        }
      }
      std::vector<int32_t> cnt(g * g, 0);
      int64_t worst = 0;
      for (int64_t v = 0; v < N; ++v)
        if (box[v] >= 0 && !is_sep[v]) worst = std::max<int64_t>(worst, ++cnt[box[v]]);
      if (worst > kMaxInterior) continue;       // refine the grid
      std::vector<int32_t> id_of_box(g * g, -1);
      interiors.clear();
      for (int64_t v = 0; v < N; ++v)
        if (box[v] >= 0 && !is_sep[v]) {
          int32_t& id = id_of_box[box[v]];
          if (id < 0) { id = (int32_t)interiors.size(); interiors.emplace_back(); }
          interior_of[v] = id;
          interiors[id].push_back((int32_t)v);
        }
      break;
    }
  }
  p.P = (int64_t)interiors.size();
  p.nI = 0;
  for (auto& s : interiors) p.nI = std::max<int64_t>(p.nI, (int64_t)s.size());
  p.nI = (p.nI + 7) / 8 * 8;
  // separator nodes (level-2 unknowns) and each interior's boundary list
  std::vector<uint8_t> keep(N);
  p.nC = 0;
  for (int64_t v = 0; v < N; ++v) {
    keep[v] = (p.u_of[v] < 0 && interior_of[v] < 0);
    p.nC += keep[v];
  }
  std::vector<std::vector<int32_t>> bnd(p.P);   // separator NODE ids adjacent to interior p
  for (int64_t ip = 0; ip < p.P; ++ip) {
    auto& b = bnd[ip];
    for (int32_t v : interiors[ip])
      for (int32_t q = p.rowptr[v]; q < p.rowptr[v + 1]; ++q) {
        const int32_t w = p.colidx[q];
        if (keep[w]) b.push_back(w);
        else if (p.u_of[w] < 0 && interior_of[w] != ip) throw std::runtime_error("plan: interiors are not separated");
      }
    std::sort(b.begin(), b.end());
    b.erase(std::unique(b.begin(), b.end()), b.end());
    p.bmax = std::max<int64_t>(p.bmax, (int64_t)b.size());
  }
  p.bmax = (p.bmax + 7) / 8 * 8;
  // ---- level-2 graph on C: mesh edges inside C plus a clique on every boundary list (the fill of
  // eliminating that interior) ---------------------------------------------------------------------
  std::vector<int32_t> rowptr2(N + 1, 0), colidx2;
  {
    std::vector<std::vector<int32_t>> adj(N);
    for (int64_t v = 0; v < N; ++v)
      if (keep[v])
        for (int32_t q = p.rowptr[v]; q < p.rowptr[v + 1]; ++q)
          if (keep[p.colidx[q]]) adj[v].push_back(p.colidx[q]);
    for (auto& b : bnd)
      for (int32_t a : b)
        for (int32_t c2 : b) adj[a].push_back(c2);
    for (int64_t v = 0; v < N; ++v) {
      auto& a = adj[v];
      std::sort(a.begin(), a.end());
      a.erase(std::unique(a.begin(), a.end()), a.end());
      colidx2.insert(colidx2.end(), a.begin(), a.end());
      rowptr2[v + 1] = (int32_t)colidx2.size();
    }
  }
  // ---- level sets of the level-2 graph (pseudo-peripheral start, restart per component) ----------
  std::vector<int32_t> committed(N, -1), stamp(N, -1);
  std::vector<std::vector<int32_t>> all_levels, lv;
  int32_t stamp_val = 0;
  for (int32_t s = 0; s < (int32_t)N; ++s) {
    if (!keep[s] || committed[s] >= 0) continue;
    bfs_levels(rowptr2, colidx2, keep, committed, stamp, stamp_val++, s, lv);
    int32_t far = lv.back()[0];
    bfs_levels(rowptr2, colidx2, keep, committed, stamp, stamp_val++, far, lv);
    far = lv.back()[0];
    bfs_levels(rowptr2, colidx2, keep, committed, stamp, stamp_val++, far, lv);
    for (auto& l : lv) {
      for (int32_t v : l) committed[v] = (int32_t)all_levels.size();
      all_levels.push_back(l);
    }
  }
  p.n_levels = (int64_t)all_levels.size();
  // ---- merge consecutive levels into blocks of >= target_block nodes ------------------------
  std::vector<std::vector<int32_t>> blocks;
  {
    std::vector<int32_t> cur;
    for (auto& l : all_levels) {
      cur.insert(cur.end(), l.begin(), l.end());
      if ((int64_t)cur.size() >= target_block) {
        blocks.push_back(cur);
        cur.clear();
      }
    }
    if (!cur.empty()) {
      if (!blocks.empty() && (int64_t)cur.size() < target_block / 2)
        blocks.back().insert(blocks.back().end(), cur.begin(), cur.end());
      else
        blocks.push_back(cur);
    }
  }
  p.nb = (int64_t)blocks.size();
  // ---- positions: [interior 0 (nI slots) | ... | interior P-1 | C in block order | U] --------------
  p.NI = p.P * p.nI;
  p.N0P = p.NI + p.nC;
  p.NP = p.N0P + p.nU;
  p.pos.assign(N, -1);
  p.node_at.assign(p.NP, -1);
  p.blk_ptr.assign(p.nb + 1, 0);
  p.blk_of.assign(p.nC, 0);
  p.max_block = 0;
  for (int64_t ip = 0; ip < p.P; ++ip)
    for (size_t a = 0; a < interiors[ip].size(); ++a) {
      p.pos[interiors[ip][a]] = (int32_t)(ip * p.nI + (int64_t)a);
      p.node_at[ip * p.nI + a] = interiors[ip][a];
    }
  {
    int32_t q = 0;
    for (int64_t k = 0; k < p.nb; ++k) {
      for (int32_t v : blocks[k]) {
        p.pos[v] = (int32_t)(p.NI + q);
        p.node_at[p.NI + q] = v;
        p.blk_of[q] = (int32_t)k;
        ++q;
      }
      p.blk_ptr[k + 1] = q;
      p.max_block = std::max<int64_t>(p.max_block, (int64_t)blocks[k].size());
    }
    if (q != p.nC) throw std::runtime_error("plan: level sets do not cover the separator set");
    for (int64_t u = 0; u < p.nU; ++u) {
      p.pos[p.U[u]] = (int32_t)(p.N0P + u);
      p.node_at[p.N0P + u] = p.U[u];
    }
  }
  p.offD.assign(p.nb + 1, 0);
  p.offS.assign(p.nb + 1, 0);
  for (int64_t k = 0; k < p.nb; ++k) {
    int64_t nk = p.blk_ptr[k + 1] - p.blk_ptr[k];
    p.offD[k + 1] = p.offD[k] + nk * nk;
    int64_t nk1 = (k + 1 < p.nb) ? p.blk_ptr[k + 2] - p.blk_ptr[k + 1] : 0;
    p.offS[k + 1] = p.offS[k] + nk1 * nk;
  }
  p.szD = p.offD[p.nb];
  p.szS = p.offS[p.nb];
  p.szPP = p.P * p.nI * p.nI;
  p.szV = p.P * p.nI * p.bmax;
  auto bsz = [&](int64_t k) { return (int64_t)(p.blk_ptr[k + 1] - p.blk_ptr[k]); };
  // offset of the separator pair (LOCAL level-2 positions) inside [diag blocks | sub-diagonal blocks]
  auto pair_off = [&](int32_t pa, int32_t pb) -> int64_t {
    int32_t ka = p.blk_of[pa], kb = p.blk_of[pb];
    int64_t la = pa - p.blk_ptr[ka], lb = pb - p.blk_ptr[kb];
    if (ka == kb) return p.offD[ka] + la * bsz(ka) + lb;
    if (ka == kb + 1) return p.szD + p.offS[kb] + la * bsz(kb) + lb;
    if (kb == ka + 1) return p.szD + p.offS[ka] + lb * bsz(ka) + la;
    throw std::runtime_error("plan: ordering is not block-tridiagonal");
  };
  // boundary lists in level-2 local positions + index of a separator node inside a boundary list
  p.bnd_loc.assign(p.P * p.bmax, -1);
  std::vector<int32_t> bcount(p.P, 0);
  for (int64_t ip = 0; ip < p.P; ++ip) {
    bcount[ip] = (int32_t)bnd[ip].size();
    for (size_t a = 0; a < bnd[ip].size(); ++a) p.bnd_loc[ip * p.bmax + a] = p.pos[bnd[ip][a]] - (int32_t)p.NI;
  }
  auto bnd_index = [&](int64_t ip, int32_t node) -> int64_t {
    auto it = std::lower_bound(bnd[ip].begin(), bnd[ip].end(), node);
    if (it == bnd[ip].end() || *it != node) throw std::runtime_error("plan: node is not on the interior's boundary");
    return (int64_t)(it - bnd[ip].begin());
  };
  // ---- Schur-complement scatter of the interiors into the level-2 blocks, as gather lists ---------
  // T_p = Bt_p^T A_p^-1 Bt_p (bmax x bmax); entry (a, b) is subtracted from the level-2 entry of
  // (bnd_p[a], bnd_p[b]) when that entry is stored (same block, or row block = column block + 1).
  {
    std::vector<std::pair<int64_t, int32_t>> items;     // (destination code, source index)
    for (int64_t ip = 0; ip < p.P; ++ip)
      for (int32_t a = 0; a < bcount[ip]; ++a)
        for (int32_t b2 = 0; b2 < bcount[ip]; ++b2) {
          const int32_t la = p.bnd_loc[ip * p.bmax + a], lb = p.bnd_loc[ip * p.bmax + b2];
          const int32_t ka = p.blk_of[la], kb = p.blk_of[lb];
          if (ka == kb || ka == kb + 1)
            items.emplace_back(pair_off(la, lb), (int32_t)(ip * p.bmax * p.bmax + a * p.bmax + b2));
          else if (kb != ka + 1)
            throw std::runtime_error("plan: level-2 ordering is not block-tridiagonal");
        }
    std::sort(items.begin(), items.end());
    p.m2_dst.clear();
    p.m2_src_ptr.assign(1, 0);
    p.m2_src.clear();
    for (size_t q = 0; q < items.size(); ++q) {
      if (q == 0 || items[q].first != items[q - 1].first) {
        if (q) p.m2_src_ptr.push_back((int32_t)p.m2_src.size());
        p.m2_dst.push_back(items[q].first);
      }
      p.m2_src.push_back(items[q].second);
    }
    if (!items.empty()) p.m2_src_ptr.push_back((int32_t)p.m2_src.size());
    // right-hand side: r_C[c] -= sum over (p, a) with bnd_p[a] == c of row a of Bt_p^T t_p
    std::vector<std::pair<int32_t, int32_t>> ritems;
    for (int64_t ip = 0; ip < p.P; ++ip)
      for (int32_t a = 0; a < bcount[ip]; ++a) ritems.emplace_back(p.bnd_loc[ip * p.bmax + a], (int32_t)(ip * p.bmax + a));
    std::sort(ritems.begin(), ritems.end());
    p.rc_ptr.assign(p.nC + 1, 0);
    p.rc_src.clear();
    for (auto& it : ritems) {
      p.rc_ptr[it.first + 1]++;
      p.rc_src.push_back(it.second);
    }
    for (int64_t c2 = 0; c2 < p.nC; ++c2) p.rc_ptr[c2 + 1] += p.rc_ptr[c2];
    // gather of the level-2 selected inverse on the boundary cliques
    p.sg_off.assign(p.P * p.bmax * p.bmax, -1);
    for (int64_t ip = 0; ip < p.P; ++ip)
      for (int32_t a = 0; a < bcount[ip]; ++a)
        for (int32_t b2 = 0; b2 < bcount[ip]; ++b2)
          p.sg_off[ip * p.bmax * p.bmax + a * p.bmax + b2] =
              pair_off(p.bnd_loc[ip * p.bmax + a], p.bnd_loc[ip * p.bmax + b2]);
  }
  // node classes: 0 interior, 1 separator, 2 electrode
  auto cls = [&](int32_t v) { return p.u_of[v] >= 0 ? 2 : (interior_of[v] >= 0 ? 0 : 1); };
  p.slot_dst.assign(p.nnz, -1);
  p.slot_kind.assign(p.nnz, 0);
  for (int64_t r = 0; r < N; ++r)
    for (int32_t s = p.rowptr[r]; s < p.rowptr[r + 1]; ++s) {
      const int32_t c2 = p.colidx[s];
      const int cr = cls((int32_t)r), cc = cls(c2);
      const int32_t pr = p.pos[r], pc = p.pos[c2];
      if (cr == 1 && cc == 1) {
        const int32_t lr = pr - (int32_t)p.NI, lc = pc - (int32_t)p.NI;
        const int32_t kr = p.blk_of[lr], kc = p.blk_of[lc];
        if (kr == kc) {
          p.slot_kind[s] = 1;
          p.slot_dst[s] = pair_off(lr, lc);
        } else if (kr == kc + 1) {
          p.slot_kind[s] = 2;
          p.slot_dst[s] = pair_off(lr, lc) - p.szD;
        } else if (kc != kr + 1) {
          throw std::runtime_error("plan: ordering is not block-tridiagonal");
        }
      } else if (cr == 0 && cc == 0) {
        const int64_t ip = interior_of[r];
        if (interior_of[c2] != ip) throw std::runtime_error("plan: interiors are not separated");
        p.slot_kind[s] = 5;
        p.slot_dst[s] = ip * p.nI * p.nI + (pr - ip * p.nI) * p.nI + (pc - ip * p.nI);
      } else if (cr == 0 && cc == 1) {
        const int64_t ip = interior_of[r];
        p.slot_kind[s] = 6;
        p.slot_dst[s] = ip * p.nI * p.bmax + (pr - ip * p.nI) * p.bmax + bnd_index(ip, c2);
      } else if (cr <= 1 && cc == 2) {
        p.slot_kind[s] = 3;
        p.slot_dst[s] = (int64_t)pr * p.nU + (pc - p.N0P);
      } else if (cr == 2 && cc == 2) {
        p.slot_kind[s] = 4;
        p.slot_dst[s] = (int64_t)(pr - p.N0P) * p.nU + (pc - p.N0P);
      }
    }
  // ---- per-element positions and selected-inverse gather offsets into
  //      [Sd | So | Sigma_pp (P nI nI) | -V (P nI bmax)] ----------------------------------------------
  p.elem_pos.resize(3 * E);
  p.g0_off.assign(6 * E, -1);
  static const int PA[6] = {0, 1, 2, 0, 1, 2}, PB[6] = {0, 1, 2, 1, 2, 0};
  const int64_t basePP = p.szD + p.szS, baseV = basePP + p.szPP;
  for (int64_t e = 0; e < E; ++e) {
    for (int a = 0; a < 3; ++a) p.elem_pos[3 * e + a] = p.pos[p.elem[3 * e + a]];
    for (int q = 0; q < 6; ++q) {
      const int32_t va = p.elem[3 * e + PA[q]], vb = p.elem[3 * e + PB[q]];
      const int ca = cls(va), cb = cls(vb);
      if (ca == 2 || cb == 2) continue;
      const int32_t pa = p.pos[va], pb = p.pos[vb];
      int64_t off;
      if (ca == 1 && cb == 1) {
        off = pair_off(pa - (int32_t)p.NI, pb - (int32_t)p.NI);
      } else if (ca == 0 && cb == 0) {
        const int64_t ip = interior_of[va];
        if (interior_of[vb] != ip) throw std::runtime_error("plan: interiors are not separated");
        off = basePP + ip * p.nI * p.nI + (pa - ip * p.nI) * p.nI + (pb - ip * p.nI);
      } else if (ca == 0) {
        const int64_t ip = interior_of[va];
        off = baseV + ip * p.nI * p.bmax + (pa - ip * p.nI) * p.bmax + bnd_index(ip, vb);
      } else {
        const int64_t ip = interior_of[vb];
        off = baseV + ip * p.nI * p.bmax + (pb - ip * p.nI) * p.bmax + bnd_index(ip, va);
      }
      p.g0_off[6 * e + q] = off;
    }
  }
  // ---- compact position space for everything downstream of the base stage (Xh, Yh, phi0, Woodbury):
  // the padding slots of the interiors are dropped; row_map sends a compact F0 row to its padded row ----
  p.ppos = p.pos;
  p.row_map.assign(p.N0, -1);
  {
    int32_t q = 0;
    for (int64_t r = 0; r < p.N0P; ++r)
      if (p.node_at[r] >= 0) {
        p.pos[p.node_at[r]] = q;
        p.row_map[q] = (int32_t)r;
        ++q;
      }
    if (q != p.N0) throw std::runtime_error("plan: position compaction lost nodes");
    for (int64_t u = 0; u < p.nU; ++u) p.pos[p.U[u]] = (int32_t)(p.N0 + u);
  }
  for (int64_t e = 0; e < E; ++e)
    for (int a = 0; a < 3; ++a) p.elem_pos[3 * e + a] = p.pos[p.elem[3 * e + a]];
}


namespace {

// ---- multi-level nested dissection (TreePlan) -------------------------------------------------------------------
// Recursive coordinate bisection of the F0 graph: a sub-domain of more than leaf_cap nodes is cut across its longer
// extent at the median coordinate (snapped to a coordinate value, so that a structured grid is cut along a grid
// line); the separator is the set of nodes of the lower half that have a neighbour in the upper half, so the two
// remaining halves are disconnected.  The reference has no ordering at all (dense inverse, CEIT/EFEM.py:111-119).
struct TreeBuilder {
  HostPlan& p;
  int64_t leaf_cap;
  std::vector<std::vector<int32_t>> own;       // per front (creation order)
  std::vector<int32_t> parent;
  std::vector<std::vector<int32_t>> children;
  std::vector<int32_t> mark;                   // node -> stamp of the upper half it belongs to
  int32_t stamp = 0;
  TreeBuilder(HostPlan& p_, int64_t cap) : p(p_), leaf_cap(cap), mark(p_.N, -1) {}

  int32_t new_front(std::vector<int32_t>&& nodes_) {
    own.push_back(std::move(nodes_));
    parent.push_back(-1);
    children.emplace_back();
    return (int32_t)own.size() - 1;
  }
  // returns the roots of the forest that covers `set`
  // One cut of `set` (destroyed) across its longer extent: (separator, lower part, upper part).
  void cut(std::vector<int32_t>& set, std::vector<int32_t>& sep, std::vector<int32_t>& lower,
           std::vector<int32_t>& upper) {
    double lo[2] = {1e300, 1e300}, hi[2] = {-1e300, -1e300};
    for (int32_t v : set)
      for (int d = 0; d < 2; ++d) {
        lo[d] = std::min(lo[d], p.nodes[2 * v + d]);
        hi[d] = std::max(hi[d], p.nodes[2 * v + d]);
      }
    const int ax = (hi[0] - lo[0] >= hi[1] - lo[1]) ? 0 : 1, ot = 1 - ax;
    const double eps = 1e-9 * std::max(hi[ax] - lo[ax], 1e-300);
    std::sort(set.begin(), set.end(), [&](int32_t a, int32_t b) {
      const double ca = p.nodes[2 * a + ax], cb = p.nodes[2 * b + ax];
      if (ca != cb) return ca < cb;
      const double oa = p.nodes[2 * a + ot], ob = p.nodes[2 * b + ot];
      if (oa != ob) return oa < ob;
      return a < b;
    });
    size_t mid = set.size() / 2;
    {
      // snap the cut to a coordinate value: every node within eps of the median coordinate goes to the upper half
      const double cm = p.nodes[2 * set[mid] + ax];
      size_t first = mid, last = mid;
      while (first > 0 && p.nodes[2 * set[first - 1] + ax] >= cm - eps) --first;
      while (last < set.size() && p.nodes[2 * set[last] + ax] <= cm + eps) ++last;
      if (first > 0) mid = first;
      else if (last < set.size()) mid = last;
      // else: every node has the same coordinate along ax (degenerate): keep the index split
    }
    const int32_t st = stamp++;
    for (size_t k = mid; k < set.size(); ++k) mark[set[k]] = st;
    upper.assign(set.begin() + mid, set.end());
    for (size_t k = 0; k < mid; ++k) {
      const int32_t v = set[k];
      bool is_cut = false;
      for (int32_t q = p.rowptr[v]; q < p.rowptr[v + 1] && !is_cut; ++q) is_cut = (mark[p.colidx[q]] == st);
      (is_cut ? sep : lower).push_back(v);
    }
    set.clear();
    set.shrink_to_fit();
  }
  // returns the roots of the forest that covers `set`.  ways = 2: one cut per front; ways = 4: both halves are cut
  // again (across the other direction on a square-ish sub-domain) and the three separators form ONE front with up to
  // four children - half as many levels on the serial spine of the base stage for a slightly larger dense block.
  int ways = 2;
  std::vector<int32_t> dissect(std::vector<int32_t>& set) {
    if (set.empty()) return {};
    if ((int64_t)set.size() <= leaf_cap) {
      std::sort(set.begin(), set.end());
      return {new_front(std::move(set))};
    }
    std::vector<int32_t> sep, lower, upper;
    cut(set, sep, lower, upper);
    std::vector<std::vector<int32_t>> parts;
    if (ways == 4) {
      for (auto* half : {&lower, &upper}) {
        if ((int64_t)half->size() > leaf_cap) {
          std::vector<int32_t> a, b;
          cut(*half, sep, a, b);                  // its separator joins this front's own set
          parts.push_back(std::move(a));
          parts.push_back(std::move(b));
        } else {
          parts.push_back(std::move(*half));
        }
      }
    } else {
      parts.push_back(std::move(lower));
      parts.push_back(std::move(upper));
    }
    std::vector<int32_t> kids;
    for (auto& part : parts) {
      std::vector<int32_t> k2 = dissect(part);
      kids.insert(kids.end(), k2.begin(), k2.end());
    }
    if (sep.empty()) return kids;                 // the parts were not connected: no front, the forest passes up
    std::sort(sep.begin(), sep.end());
    const int32_t f = new_front(std::move(sep));
    for (int32_t c : kids) parent[c] = f;
    children[f] = kids;
    return {f};
  }
};

void build_tree_plan(HostPlan& p, int64_t leaf_cap, int pull_mode, int ways, int parts, int part) {
  const int64_t N = p.N, E = p.E;
  if (parts > 1 && !pull_mode) throw std::runtime_error("plan: the domain decomposition needs the pull schedule (nd_pull = 1)");
  TreeBuilder tb(p, leaf_cap);
  tb.ways = (ways == 4) ? 4 : 2;
  {
    std::vector<int32_t> all;
    all.reserve(p.N0);
    for (int64_t v = 0; v < N; ++v)
      if (p.u_of[v] < 0) all.push_back((int32_t)v);
    tb.dissect(all);
  }
  const int32_t nF = (int32_t)tb.own.size();
  // heights -> levels; global front ids level by level (creation order inside a level)
  std::vector<int32_t> height(nF, 0);
  for (int32_t f = 0; f < nF; ++f)            // children are created before their parent
    for (int32_t c : tb.children[f]) height[f] = std::max(height[f], height[c] + 1);
  int32_t Hh = 0;
  for (int32_t f = 0; f < nF; ++f) Hh = std::max(Hh, height[f] + 1);
  // fronts with a non-zero right-hand side K_0U: a node next to an electrode node, or an active child
  std::vector<uint8_t> active(nF, 0);
  for (int32_t f = 0; f < nF; ++f) {          // children are created before their parent
    for (int32_t v : tb.own[f])
      for (int32_t q = p.rowptr[v]; q < p.rowptr[v + 1] && !active[f]; ++q) active[f] = (p.u_of[p.colidx[q]] >= 0);
    for (int32_t c : tb.children[f]) active[f] |= active[c];
  }
  // ---- domain decomposition: cut the forest at the smallest depth that has >= parts fronts; the sub-trees rooted
  // there go to the ranks in contiguous groups of creation order (balanced by node count), everything above is the
  // shared top (frank = -1) ----------------------------------------------------------------------------------------
  std::vector<int32_t> frank(nF, -1);
  int32_t dcut = 0;
  if (parts > 1) {
    std::vector<int32_t> depth(nF, 0);
    for (int32_t f = nF - 1; f >= 0; --f) depth[f] = tb.parent[f] < 0 ? 0 : depth[tb.parent[f]] + 1;   // parents have larger ids
    int32_t dmax = 0;
    for (int32_t f = 0; f < nF; ++f) dmax = std::max(dmax, depth[f]);
    std::vector<int32_t> cnt(dmax + 1, 0);
    for (int32_t f = 0; f < nF; ++f) cnt[depth[f]]++;
    dcut = -1;
    for (int32_t d = 0; d <= dmax && dcut < 0; ++d)
      if (cnt[d] >= parts) dcut = d;
    if (dcut < 0) throw std::runtime_error("plan: the elimination forest has fewer sub-trees than ranks");
    std::vector<int64_t> weight(nF, 0);
    for (int32_t f = 0; f < nF; ++f) {        // children first
      weight[f] += (int64_t)tb.own[f].size();
      if (tb.parent[f] >= 0) weight[tb.parent[f]] += weight[f];
    }
    int64_t total = 0;
    for (int32_t f = 0; f < nF; ++f)
      if (depth[f] == dcut) total += weight[f];
    int64_t cum = 0;
    for (int32_t f = 0; f < nF; ++f)
      if (depth[f] == dcut) {
        frank[f] = (int32_t)std::min<int64_t>(parts - 1, ((2 * cum + weight[f]) * parts) / (2 * std::max<int64_t>(total, 1)));
        cum += weight[f];
      }
    for (int32_t f = nF - 1; f >= 0; --f)
      if (depth[f] > dcut) frank[f] = frank[tb.parent[f]];
  }
  // segments: (height, owner) classes; the rank-owned ones first, ordered by (height, owner), then the top by height
  std::vector<int32_t> gid(nF, -1), old_of(nF, -1);
  TreePlan& t = p.tree;
  t = TreePlan();
  t.nF = nF;
  t.parts = parts;
  t.part = part;
  t.dcut = dcut;
  std::vector<int32_t> seg_of_old(nF, -1);
  {
    int32_t next = 0;
    for (int top = 0; top < 2; ++top)
      for (int32_t l = 0; l < Hh; ++l)
        for (int32_t r = (top ? -1 : 0); r < (top ? 0 : parts); ++r) {
          TreeLevel lv;
          lv.f0 = next;
          lv.owner = r;
          lv.height = l;
          for (int pass = 1; pass >= 0; --pass) {   // active fronts first
            for (int32_t f = 0; f < nF; ++f)
              if (height[f] == l && frank[f] == r && active[f] == pass) {
                gid[f] = next;
                old_of[next] = f;
                seg_of_old[f] = (int32_t)t.lv.size();
                ++next;
              }
            if (pass == 1) lv.na = next - lv.f0;
          }
          lv.n = next - lv.f0;
          if (lv.n == 0) continue;
          if (!top) t.n_own_segs = (int32_t)t.lv.size() + 1;
          t.lv.push_back(lv);
        }
    if (next != nF) throw std::runtime_error("plan: segment numbering lost fronts");
  }
  const int32_t H = (int32_t)t.lv.size();
  t.H = H;
  // owner / slot of every F0 node, Euler intervals for ancestor tests
  std::vector<int32_t> owner(N, -1), slot(N, -1), level_of(nF, 0), seg_of(nF, 0);
  for (int32_t g = 0; g < nF; ++g) {
    const auto& o = tb.own[old_of[g]];
    level_of[g] = height[old_of[g]];            // height: ordering / ancestor tests
    seg_of[g] = seg_of_old[old_of[g]];          // segment: storage
    for (size_t a = 0; a < o.size(); ++a) {
      owner[o[a]] = g;
      slot[o[a]] = (int32_t)a;
    }
  }
  std::vector<int32_t> par(nF, -1);
  for (int32_t g = 0; g < nF; ++g) par[g] = tb.parent[old_of[g]] < 0 ? -1 : gid[tb.parent[old_of[g]]];
  auto is_proper_ancestor = [&](int32_t a, int32_t f) {     // a is a proper ancestor of f (levels strictly increase)
    int32_t x = par[f];
    while (x >= 0 && level_of[x] < level_of[a]) x = par[x];
    return x == a;
  };
  // boundary lists bottom-up (sorted by node id)
  std::vector<std::vector<int32_t>> bnd(nF);
  for (int32_t g = 0; g < nF; ++g) {                        // level order: children before parents
    auto& b = bnd[g];
    for (int32_t v : tb.own[old_of[g]])
      for (int32_t q = p.rowptr[v]; q < p.rowptr[v + 1]; ++q) {
        const int32_t w = p.colidx[q];
        if (p.u_of[w] >= 0 || owner[w] == g) continue;
        if (level_of[owner[w]] > level_of[g]) {
          if (!is_proper_ancestor(owner[w], g)) throw std::runtime_error("plan: sub-domains are not separated");
          b.push_back(w);
        }
        // else: w belongs to a descendant, which lists v in its own boundary
      }
    for (int32_t c : tb.children[old_of[g]])
      for (int32_t w : bnd[gid[c]])
        if (owner[w] != g) b.push_back(w);
    std::sort(b.begin(), b.end());
    b.erase(std::unique(b.begin(), b.end()), b.end());
    if (par[g] < 0 && !b.empty()) throw std::runtime_error("plan: a root front has a boundary");
  }
  // level sizes and offsets
  p.N0P = 0;
  for (int32_t l = 0; l < H; ++l) {
    TreeLevel& lv = t.lv[l];
    int32_t s = 0, b = 0;
    for (int32_t k = 0; k < lv.n; ++k) {
      s = std::max<int32_t>(s, (int32_t)tb.own[old_of[lv.f0 + k]].size());
      b = std::max<int32_t>(b, (int32_t)bnd[lv.f0 + k].size());
    }
    lv.s = (s + 7) / 8 * 8;
    lv.b = (b + 7) / 8 * 8;
    lv.offA = t.szA;
    lv.offB = t.szB;
    lv.offC = t.szC;
    lv.offR = (int64_t)t.rel.size();
    t.szA += (int64_t)lv.n * lv.s * lv.s;
    t.szB += (int64_t)lv.n * lv.s * lv.b;
    t.szC += (int64_t)lv.n * lv.b * lv.b;
    t.maxG = std::max<int64_t>(t.maxG, (int64_t)lv.n * lv.b);
    t.maxC = std::max<int64_t>(t.maxC, (int64_t)lv.n * lv.b * lv.b);
    t.rel.resize(t.rel.size() + (size_t)lv.n * lv.b, -1);
  }
  {
    // padded rows: the active fronts rank by rank (each rank's segments by height), then the active fronts of the top,
    // then the inactive ones - so that S_U and the sweeps over non-zero right-hand sides touch [0, rowsA) only and a
    // rank's share of them is one contiguous range
    t.ra0.assign(parts, 0);
    t.ra1.assign(parts, 0);
    for (int32_t r = 0; r <= parts; ++r) {
      const int32_t who = (r == parts) ? -1 : r;
      if (who >= 0) t.ra0[who] = p.N0P;
      else t.ra_top0 = p.N0P;
      for (int32_t l = 0; l < H; ++l)
        if (t.lv[l].owner == who) {
          t.lv[l].row0 = p.N0P;
          p.N0P += (int64_t)t.lv[l].na * t.lv[l].s;
        }
      if (who >= 0) t.ra1[who] = p.N0P;
    }
    t.rowsA = p.N0P;
    for (int32_t l = 0; l < H; ++l) {
      t.lv[l].row0i = p.N0P;
      p.N0P += (int64_t)(t.lv[l].n - t.lv[l].na) * t.lv[l].s;
    }
  }
  if (p.N0P > (int64_t)2000000000) throw std::runtime_error("plan: padded row space exceeds 32-bit indices");
  t.bnd_row.assign(t.rel.size(), -1);
  // positions
  p.P = p.nI = p.bmax = p.nC = p.NI = p.nb = p.max_block = p.n_levels = 0;
  p.szD = p.szS = 0;
  p.szPP = t.szA;
  p.szV = t.szB;
  p.NP = p.N0P + p.nU;
  p.pos.assign(N, -1);
  p.node_at.assign(p.NP, -1);
  auto prow = [&](int32_t v) -> int64_t {
    const int32_t g = owner[v];
    const TreeLevel& lv = t.lv[seg_of[g]];
    const int64_t k = g - lv.f0;
    return (k < lv.na ? lv.row0 + k * lv.s : lv.row0i + (k - lv.na) * lv.s) + slot[v];
  };
  for (int64_t v = 0; v < N; ++v)
    if (p.u_of[v] < 0) {
      p.pos[v] = (int32_t)prow((int32_t)v);
      p.node_at[p.pos[v]] = (int32_t)v;
    }
  for (int64_t u = 0; u < p.nU; ++u) {
    p.pos[p.U[u]] = (int32_t)(p.N0P + u);
    p.node_at[p.N0P + u] = p.U[u];
  }
  for (const TreeLevel& lv : t.lv)
    for (int64_t f = 0; f < lv.n; ++f) {
      const int64_t r0 = f < lv.na ? lv.row0 + f * lv.s : lv.row0i + (f - lv.na) * lv.s;
      for (int64_t a = 0; a < lv.s; ++a)
        if (p.node_at[r0 + a] < 0) t.pad_diag.push_back(lv.offA + f * lv.s * lv.s + a * lv.s + a);
    }
  auto bidx = [&](int32_t g, int32_t node) -> int64_t {
    auto it = std::lower_bound(bnd[g].begin(), bnd[g].end(), node);
    if (it == bnd[g].end() || *it != node) throw std::runtime_error("plan: node is not on the front's boundary");
    return (int64_t)(it - bnd[g].begin());
  };
  auto blockA = [&](int32_t g) { const TreeLevel& lv = t.lv[seg_of[g]]; return lv.offA + (int64_t)(g - lv.f0) * lv.s * lv.s; };
  auto blockB = [&](int32_t g) { const TreeLevel& lv = t.lv[seg_of[g]]; return lv.offB + (int64_t)(g - lv.f0) * lv.s * lv.b; };
  auto blockC = [&](int32_t g) { const TreeLevel& lv = t.lv[seg_of[g]]; return lv.offC + (int64_t)(g - lv.f0) * lv.b * lv.b; };
  // per-front parent data, relative indices, boundary rows, sibling index
  t.parent = par;
  t.sib.assign(nF, 0);
  t.pA.assign(nF, -1);
  t.pB.assign(nF, -1);
  t.pC.assign(nF, -1);
  t.ps.assign(nF, 0);
  t.pb.assign(nF, 0);
  for (int32_t g = 0; g < nF; ++g) {
    const auto& kids = tb.children[old_of[g]];
    for (size_t k = 0; k < kids.size(); ++k) t.sib[gid[kids[k]]] = (int32_t)k;
  }
  for (int32_t l = 0; l < H; ++l) {
    TreeLevel& lv = t.lv[l];
    for (int32_t k = 0; k < lv.n; ++k) {
      const int32_t g = lv.f0 + k, pg = par[g];
      lv.passes = std::max(lv.passes, t.sib[g] + 1);
      if (pg >= 0) {
        const TreeLevel& pl = t.lv[seg_of[pg]];
        t.pA[g] = blockA(pg);
        t.pB[g] = blockB(pg);
        t.pC[g] = blockC(pg);
        t.ps[g] = pl.s;
        t.pb[g] = pl.b;
      }
      for (size_t a = 0; a < bnd[g].size(); ++a) {
        const int32_t w = bnd[g][a];
        const int64_t q = lv.offR + (int64_t)k * lv.b + (int64_t)a;
        t.bnd_row[q] = p.pos[w];
        t.rel[q] = (owner[w] == pg) ? slot[w] : (int32_t)(t.ps[g] + bidx(pg, w));
      }
    }
  }
  // pull levels (small fronts) and the child lists they read; children of a pulling parent do not push
  // pull_mode 0: children push (extend-add passes at the child's level); 1: every level pulls (one launch per level,
  // any number of children); levels of small fronts (s, b <= 64) may additionally run the fused front kernel
  for (int32_t l = 0; l < H; ++l) {
    t.lv[l].pull = pull_mode ? 1 : 0;
    t.lv[l].fusable = (pull_mode && t.lv[l].s <= 64 && t.lv[l].b <= 64) ? 1 : 0;
  }
  t.child_ptr.assign(nF + 1, 0);
  t.f_b.assign(nF, 0);
  t.f_offC.assign(nF, 0);
  t.f_relOff.assign(nF, 0);
  t.push.assign(nF, 0);
  for (int32_t g = 0; g < nF; ++g) t.child_ptr[g + 1] = t.child_ptr[g] + (int32_t)tb.children[old_of[g]].size();
  t.child_ids.resize(t.child_ptr[nF]);
  for (int32_t g = 0; g < nF; ++g) {
    const auto& kids = tb.children[old_of[g]];
    for (size_t k = 0; k < kids.size(); ++k) t.child_ids[t.child_ptr[g] + (int32_t)k] = gid[kids[k]];
    const TreeLevel& lv = t.lv[seg_of[g]];
    t.f_b[g] = lv.b;
    t.f_offC[g] = blockC(g);
    t.f_relOff[g] = lv.offR + (int64_t)(g - lv.f0) * lv.b;
    t.push[g] = (par[g] >= 0 && !t.lv[seg_of[par[g]]].pull) ? 1 : 0;
  }
  // inverse relative indices for the destination-centric pull: for child ch, position q of the PARENT's front index
  // space [0, ps + pb) -> boundary slot of ch that maps there, or -1
  t.f_invOff.assign(nF, 0);
  if (pull_mode) {
    for (int32_t g = 0; g < nF; ++g) {
      t.f_invOff[g] = (int64_t)t.inv_rel.size();
      if (par[g] < 0) continue;
      const int64_t span = (int64_t)t.ps[g] + t.pb[g];
      t.inv_rel.resize(t.inv_rel.size() + span, -1);
      const TreeLevel& lv = t.lv[seg_of[g]];
      for (int32_t a = 0; a < lv.b; ++a) {
        const int32_t r = t.rel[t.f_relOff[g] + a];
        if (r >= 0) t.inv_rel[t.f_invOff[g] + r] = a;
      }
    }
  }
  for (int32_t l = 0; l < H; ++l) {               // extend-add passes of a level: only for the fronts that push
    TreeLevel& lv = t.lv[l];
    lv.passes = 0;
    for (int32_t k = 0; k < lv.n; ++k)
      if (t.push[lv.f0 + k]) lv.passes = std::max(lv.passes, t.sib[lv.f0 + k] + 1);
  }
  // right-hand-side scatter lists per level (destination-major, deterministic source order)
  t.rc_ptr.assign(1, 0);
  for (int32_t l = 0; l < H; ++l) {
    TreeLevel& lv = t.lv[l];
    std::vector<std::pair<int32_t, int32_t>> items;
    for (int32_t k = 0; k < lv.na; ++k)           // inactive fronts have a zero right-hand side: no contribution
      for (size_t a = 0; a < bnd[lv.f0 + k].size(); ++a)
        items.emplace_back(p.pos[bnd[lv.f0 + k][a]], (int32_t)(k * lv.b + (int32_t)a));
    std::sort(items.begin(), items.end());
    lv.rc0 = (int64_t)t.rc_dst.size();
    for (size_t q = 0; q < items.size(); ++q) {
      if (q == 0 || items[q].first != items[q - 1].first) {
        if (q) t.rc_ptr.push_back((int32_t)t.rc_src.size());
        t.rc_dst.push_back(items[q].first);
      }
      t.rc_src.push_back(items[q].second);
    }
    if (!items.empty()) t.rc_ptr.push_back((int32_t)t.rc_src.size());
    lv.rc_n = (int64_t)t.rc_dst.size() - lv.rc0;
  }
  // ---- CSR slot -> dense destination: kind 5 own block (A storage), 6 coupling (B storage, the row's owner is the
  // descendant), 3 K0U, 4 KUU; the mirrored entry of a coupling pair is not stored -----------------------------------
  p.slot_dst.assign(p.nnz, -1);
  p.slot_kind.assign(p.nnz, 0);
  for (int64_t r = 0; r < N; ++r)
    for (int32_t s = p.rowptr[r]; s < p.rowptr[r + 1]; ++s) {
      const int32_t c2 = p.colidx[s];
      const bool ru = p.u_of[r] >= 0, cu = p.u_of[c2] >= 0;
      if (!ru && !cu) {
        const int32_t gr = owner[r], gc = owner[c2];
        if (gr == gc) {
          p.slot_kind[s] = 5;
          p.slot_dst[s] = blockA(gr) + (int64_t)slot[r] * t.lv[seg_of[gr]].s + slot[c2];
        } else if (level_of[gr] < level_of[gc]) {
          p.slot_kind[s] = 6;
          p.slot_dst[s] = blockB(gr) + (int64_t)slot[r] * t.lv[seg_of[gr]].b + bidx(gr, c2);
        }
      } else if (!ru && cu) {
        if (p.pos[r] >= t.rowsA) throw std::runtime_error("plan: a right-hand-side row lies outside the active region");
        p.slot_kind[s] = 3;
        p.slot_dst[s] = (int64_t)p.pos[r] * p.nU + p.u_of[c2];
      } else if (ru && cu) {
        p.slot_kind[s] = 4;
        p.slot_dst[s] = (int64_t)p.u_of[r] * p.nU + p.u_of[c2];
      }
    }
  // ---- per-element gather offsets into the selected inverse [Sigma_own (A storage) | V (B storage)] ---------------
  p.g0_off.assign(6 * E, -1);
  static const int PA[6] = {0, 1, 2, 0, 1, 2}, PB[6] = {0, 1, 2, 1, 2, 0};
  for (int64_t e = 0; e < E; ++e)
    for (int q = 0; q < 6; ++q) {
      const int32_t va = p.elem[3 * e + PA[q]], vb = p.elem[3 * e + PB[q]];
      if (p.u_of[va] >= 0 || p.u_of[vb] >= 0) continue;
      const int32_t ga = owner[va], gb = owner[vb];
      if (ga == gb) p.g0_off[6 * e + q] = blockA(ga) + (int64_t)slot[va] * t.lv[seg_of[ga]].s + slot[vb];
      else if (level_of[ga] < level_of[gb])
        p.g0_off[6 * e + q] = t.szA + blockB(ga) + (int64_t)slot[va] * t.lv[seg_of[ga]].b + bidx(ga, vb);
      else
        p.g0_off[6 * e + q] = t.szA + blockB(gb) + (int64_t)slot[vb] * t.lv[seg_of[gb]].b + bidx(gb, va);
    }
  // ---- domain decomposition: export slots of the sub-tree roots, element ownership ----------------------------------
  auto rank_of_front = [&](int32_t g) { return t.lv[seg_of[g]].owner; };
  if (parts > 1) {
    std::vector<std::vector<int32_t>> roots(parts);
    for (int32_t g = 0; g < nF; ++g) {
      const int32_t r = rank_of_front(g);
      if (r >= 0 && par[g] >= 0 && rank_of_front(par[g]) < 0) roots[r].push_back(g);
    }
    t.roots_per_rank = 0;
    t.slot_sz = 0;
    for (int32_t r = 0; r < parts; ++r) {
      t.roots_per_rank = std::max<int32_t>(t.roots_per_rank, (int32_t)roots[r].size());
      for (int32_t g : roots[r]) t.slot_sz = std::max<int64_t>(t.slot_sz, (int64_t)t.f_b[g] * t.f_b[g]);
    }
    t.exp_b.assign((size_t)parts * t.roots_per_rank, 0);
    t.exp_offC.assign((size_t)parts * t.roots_per_rank, 0);
    for (int32_t r = 0; r < parts; ++r)
      for (size_t k = 0; k < roots[r].size(); ++k) {
        t.exp_b[(size_t)r * t.roots_per_rank + k] = t.f_b[roots[r][k]];
        t.exp_offC[(size_t)r * t.roots_per_rank + k] = t.f_offC[roots[r][k]];
      }
    // an element belongs to the rank whose sub-trees hold one of its nodes (its other nodes are then in the same
    // rank's sub-trees, in the top or on an electrode: sub-domains of different ranks are separated by the top);
    // elements that touch only the top / electrodes are dealt round-robin
    for (int64_t e = 0; e < E; ++e) {
      int32_t r = -1;
      for (int a = 0; a < 3; ++a) {
        const int32_t v = p.elem[3 * e + a];
        if (p.u_of[v] >= 0) continue;
        const int32_t rv = rank_of_front(owner[v]);
        if (rv < 0) continue;
        if (r >= 0 && rv != r) throw std::runtime_error("plan: an element spans the regions of two ranks");
        r = rv;
      }
      if (r < 0) r = (int32_t)(e % parts);
      p.elem_own[e] = (r == part) ? 1 : 0;
      if (!p.elem_own[e])
        for (int q = 0; q < 6; ++q) p.g0_off[6 * e + q] = -1;
    }
  }
  // ---- compact position space (padding slots dropped; a domain-decomposed plan keeps only the rows of this rank's
  // segments and of the top) --------------------------------------------------------------------------------------------
  p.ppos = p.pos;
  {
    std::vector<int32_t> rows;
    for (int64_t r = 0; r < p.N0P; ++r) {
      const int32_t v = p.node_at[r];
      if (v < 0) continue;
      const int32_t rk = rank_of_front(owner[v]);
      if (rk >= 0 && rk != part) {
        p.pos[v] = -1;                               // foreign region: no row on this rank
        continue;
      }
      p.pos[v] = (int32_t)rows.size();
      rows.push_back((int32_t)r);
    }
    p.row_map = rows;
    p.N0c = (int64_t)rows.size();
    p.Nc = p.N0c + p.nU;
    if (parts == 1 && p.N0c != p.N0) throw std::runtime_error("plan: position compaction lost nodes");
    for (int64_t u = 0; u < p.nU; ++u) p.pos[p.U[u]] = (int32_t)(p.N0c + u);
  }
  p.elem_pos.resize(3 * E);
  for (int64_t e = 0; e < E; ++e)
    for (int a = 0; a < 3; ++a) p.elem_pos[3 * e + a] = p.pos[p.elem[3 * e + a]];
  if (parts > 1)
    for (int64_t e = 0; e < E; ++e)
      if (p.elem_own[e])
        for (int a = 0; a < 3; ++a)
          if (p.elem_pos[3 * e + a] < 0) throw std::runtime_error("plan: an owned element has a node outside the rank's rows");
  p.blk_ptr.assign(1, 0);
  p.offD.assign(1, 0);
  p.offS.assign(1, 0);
  p.use_tree = true;
  p.n_levels = H;
}

}  // namespace

// Greedy clustering: grow a patch breadth-first over the element adjacency (elements sharing a node),
// always taking a queued element if it fits the node cap; close the patch when nothing queued fits.
void build_patches(HostPlan& p, int32_t node_cap, int32_t elem_cap) {
  if (node_cap > 255) node_cap = 255;
  if (node_cap < 3) node_cap = 3;
  const int64_t N = p.N, E = p.E;
  p.patch_node_cap = node_cap;
  // node -> elements
  std::vector<int32_t> nptr(N + 1, 0), nel(3 * E);
  for (int64_t q = 0; q < 3 * E; ++q) nptr[p.elem[q] + 1]++;
  for (int64_t v = 0; v < N; ++v) nptr[v + 1] += nptr[v];
  {
    std::vector<int32_t> fill(nptr.begin(), nptr.end() - 1);
    for (int64_t e = 0; e < E; ++e)
      for (int a = 0; a < 3; ++a) nel[fill[p.elem[3 * e + a]]++] = (int32_t)e;
  }
  std::vector<uint8_t> done(E, 0), queued(E, 0);
  if ((int64_t)p.elem_own.size() == E)
    for (int64_t e = 0; e < E; ++e) done[e] = p.elem_own[e] ? 0 : 1;   // domain-decomposed plan: own elements only
  std::vector<int32_t> local_of(N, -1);
  p.patch_eptr.assign(1, 0);
  p.patch_nptr.assign(1, 0);
  p.patch_elems.clear();
  p.patch_nodes.clear();
  p.patch_local.clear();
  p.patch_elems.reserve(E);
  p.patch_local.reserve(3 * E);
  std::vector<int32_t> queue, cur_nodes, leftovers;
  for (int64_t seed = 0; seed < E; ++seed) {
    if (done[seed]) continue;
    queue.clear();
    cur_nodes.clear();
    leftovers.clear();
    queue.push_back((int32_t)seed);
    queued[seed] = 1;
    size_t head = 0;
    int32_t n_in_patch = 0;
    while (head < queue.size() && n_in_patch < elem_cap) {
      const int32_t e = queue[head++];
      int add = 0;
      for (int a = 0; a < 3; ++a)
        if (local_of[p.elem[3 * e + a]] < 0) ++add;
      // count distinct new nodes (a degenerate element could repeat a node; meshes here do not)
      if ((int32_t)cur_nodes.size() + add > node_cap) {
        leftovers.push_back(e);
        continue;
      }
      done[e] = 1;
      ++n_in_patch;
      p.patch_elems.push_back(e);
      for (int a = 0; a < 3; ++a) {
        const int32_t v = p.elem[3 * e + a];
        if (local_of[v] < 0) {
          local_of[v] = (int32_t)cur_nodes.size();
          cur_nodes.push_back(v);
        }
        p.patch_local.push_back((uint8_t)local_of[v]);
      }
      for (int a = 0; a < 3; ++a) {
        const int32_t v = p.elem[3 * e + a];
        for (int32_t q = nptr[v]; q < nptr[v + 1]; ++q) {
          const int32_t f = nel[q];
          if (!done[f] && !queued[f]) {
            queued[f] = 1;
            queue.push_back(f);
          }
        }
      }
    }
    for (int32_t e : queue)
      if (!done[e]) queued[e] = 0;
    // Shared-memory slot of every patch node.  The Woodbury kernel reads, per 16-byte load, one staged row per lane
    // (lane = element, one load per corner); the 8 lanes of a quarter warp are conflict-free when their DISTINCT
    // rows have distinct slot numbers mod 8 (woodbury_lr_pitch).  Greedy: nodes that appear in the most
    // (quarter, corner) groups first, each takes the free slot whose residue collides least inside its groups.
    {
      const int32_t nn = (int32_t)cur_nodes.size();
      const size_t l_beg = 3 * (size_t)p.patch_eptr.back();
      const int32_t ne = (int32_t)(p.patch_elems.size() - (size_t)p.patch_eptr.back());
      const int32_t n_groups = 3 * ((ne + 7) / 8);
      std::vector<std::vector<int32_t>> groups_of(nn);
      for (int32_t l = 0; l < ne; ++l)
        for (int a = 0; a < 3; ++a) {
          const int32_t v = p.patch_local[l_beg + 3 * l + a], g = 3 * (l / 8) + a;
          if (std::find(groups_of[v].begin(), groups_of[v].end(), g) == groups_of[v].end()) groups_of[v].push_back(g);
        }
      std::vector<int32_t> order(nn), slot_of(nn, -1), cnt((size_t)n_groups * 8, 0);
      for (int32_t v = 0; v < nn; ++v) order[v] = v;
      std::stable_sort(order.begin(), order.end(),
                       [&](int32_t a, int32_t b) { return groups_of[a].size() > groups_of[b].size(); });
      std::vector<uint8_t> used(nn, 0);
      for (int32_t v : order) {
        int32_t best = -1, best_cost = 1 << 30;
        for (int32_t sl = 0; sl < nn; ++sl) {
          if (used[sl]) continue;
          int32_t cost = 0;
          for (int32_t g : groups_of[v]) cost += cnt[(size_t)g * 8 + (sl & 7)];
          if (cost < best_cost) {
            best_cost = cost;
            best = sl;
          }
        }
        used[best] = 1;
        slot_of[v] = best;
        for (int32_t g : groups_of[v]) cnt[(size_t)g * 8 + (best & 7)]++;
      }
      // refinement: swap the slots of a node that still collides with any other node while that lowers the number
      // of collisions (cost of a group = sum over residues of max(count - 1, 0))
      auto group_cost = [&](int32_t g) {
        int32_t c = 0;
        for (int r = 0; r < 8; ++r) c += std::max(cnt[(size_t)g * 8 + r] - 1, 0);
        return c;
      };
      auto move = [&](int32_t v, int32_t from, int32_t to) {
        for (int32_t g : groups_of[v]) {
          cnt[(size_t)g * 8 + from]--;
          cnt[(size_t)g * 8 + to]++;
        }
      };
      std::vector<int32_t> touched;
      for (int pass = 0; pass < 4; ++pass) {
        bool improved = false;
        for (int32_t u = 0; u < nn; ++u) {
          bool collides = false;
          for (int32_t g : groups_of[u]) collides |= cnt[(size_t)g * 8 + (slot_of[u] & 7)] > 1;
          if (!collides) continue;
          for (int32_t v = 0; v < nn; ++v) {
            const int32_t ru = slot_of[u] & 7, rv = slot_of[v] & 7;
            if (v == u || ru == rv) continue;
            touched = groups_of[u];
            for (int32_t g : groups_of[v])
              if (std::find(touched.begin(), touched.end(), g) == touched.end()) touched.push_back(g);
            int32_t before = 0, after = 0;
            for (int32_t g : touched) before += group_cost(g);
            move(u, ru, rv);
            move(v, rv, ru);
            for (int32_t g : touched) after += group_cost(g);
            if (after < before) {
              std::swap(slot_of[u], slot_of[v]);
              improved = true;
              break;                      // u has a new slot: re-examine it on the next pass
            }
            move(u, rv, ru);
            move(v, ru, rv);
          }
        }
        if (!improved) break;
      }
      std::vector<int32_t> by_slot(nn);
      for (int32_t v = 0; v < nn; ++v) by_slot[slot_of[v]] = cur_nodes[v];
      for (size_t q = l_beg; q < p.patch_local.size(); ++q) p.patch_local[q] = (uint8_t)slot_of[p.patch_local[q]];
      for (int32_t v : cur_nodes) local_of[v] = -1;
      cur_nodes = by_slot;
    }
    for (int32_t v : cur_nodes) p.patch_nodes.push_back(p.pos[v]);
    p.patch_eptr.push_back((int32_t)p.patch_elems.size());
    p.patch_nptr.push_back((int32_t)p.patch_nodes.size());
  }
}

}  // namespace ceit
