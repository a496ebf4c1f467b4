// Host-side symbolic plan (pure C++, no CUDA): everything about a mesh + electrode layout that
// does not depend on material values.  See include/ceit_b200.h and DESIGN.md §3.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace ceit {

// Multi-level nested dissection of F0 (the non-electrode nodes): an elimination FOREST of fronts.  A leaf front owns
// the <= leaf_cap nodes of one final sub-domain, an internal front owns the vertex separator that split its
// sub-domain.  bnd(f) = the nodes of proper ancestors that are adjacent to f's sub-domain once it is eliminated.
// Fronts of equal height form a LEVEL and are processed as one batch (own / boundary sizes padded to the level's
// maximum s / b).  Front index space of f: [0, s) own slots, [s, s + b) boundary slots.
struct TreeLevel {
  int32_t n = 0, s = 0, b = 0;       // fronts, padded own size, padded boundary size (b = 0: roots only)
  int32_t passes = 0;                // 1 + largest sibling index (extend-add passes: siblings hit the same parent)
  int32_t f0 = 0;                    // global id of the level's first front
  int32_t na = 0;                    // fronts with a non-zero right-hand side (they come first in the level): a front
                                     // is ACTIVE when one of its nodes neighbours an electrode node or a child is active
  int64_t row0 = 0;                  // first padded row of the active fronts (front k < na: row0 + k s ...)
  int64_t row0i = 0;                 // first padded row of the inactive fronts (front k >= na: row0i + (k - na) s ...)
  int64_t offA = 0, offB = 0, offC = 0;   // element offsets of the level in the (n s s) / (n s b) / (n b b) storages
  int64_t offR = 0;                  // offset of the level in rel / bnd_row (n b entries)
  int64_t rc0 = 0, rc_n = 0;         // destinations of the level's right-hand-side scatter: rc_dst[rc0 .. rc0 + rc_n)
  int32_t pull = 0;                  // 1: the fronts of this level PULL their children's update matrices (one launch per
                                     // level whatever the number of children; children in fixed order: deterministic);
                                     // 0: the children push (one extend-add pass per sibling index at the child's level)
  int32_t fusable = 0;               // pull level of small fronts (s <= 64, b <= 64): pull + factorisation + Y / U /
                                     // A^-1 / W may run as ONE kernel per level (front_fused_kernel)
  int32_t owner = -1;                // domain-decomposed plans (TreePlan::parts > 1): rank that owns the fronts of this
                                     // SEGMENT (fronts of one height inside one rank's sub-trees); -1 = the shared top of
                                     // the forest, processed by every rank.  A rank executes owner == part and owner == -1.
  int32_t height = 0;                // height of the segment's fronts in the forest (0 = leaves)
};
struct TreePlan {
  int32_t H = 0;                     // segments (= levels when parts == 1), children before parents: the rank-owned
                                     // segments [0, n_own_segs) ordered by (height, owner), then the shared top by height
  // ---- domain decomposition over `parts` ranks (multi-GPU base stage) --------------------------
  // The forest is cut at depth `dcut`: the sub-trees below belong to one rank each (contiguous groups in creation
  // order = spatially compact regions), the fronts above are the TOP and are replicated.  One exchange between the
  // two halves of the bottom-up pass: the update matrices of the sub-tree roots are all-gathered (`exp_*`: one slot of
  // slot_sz elements per root, roots_per_rank slots per rank), the right-hand-side rows of the top and the partial
  // S_U products are all-reduced; everything after that is local to the rank's own region.
  int32_t parts = 1, part = 0, dcut = 0;
  int32_t n_own_segs = 0;
  std::vector<int64_t> ra0, ra1;     // per rank: padded rows [ra0, ra1) = the active fronts of its segments
  int64_t ra_top0 = 0;               // padded rows [ra_top0, rowsA) = the active fronts of the top
  int32_t roots_per_rank = 0;
  int64_t slot_sz = 0;               // elements per export slot (max b^2 over the sub-tree roots)
  std::vector<int32_t> exp_b;        // parts * roots_per_rank: padded boundary size of the slot's front (0 = empty slot)
  std::vector<int64_t> exp_offC;     // offset of its update matrix in the C storage
  int64_t nF = 0, szA = 0, szB = 0, szC = 0, maxG = 0, maxC = 0;   // maxG = max n b, maxC = max n b b over levels
  int64_t rowsA = 0;                 // padded rows [0, rowsA) = the active fronts of every level (K0U is zero elsewhere)
  std::vector<TreeLevel> lv;
  // per front (global id, level by level)
  std::vector<int32_t> parent, sib;  // parent front (-1: root), index among the parent's children
  std::vector<int64_t> pA, pB, pC;   // element offsets of the PARENT's own / coupling / boundary blocks
  std::vector<int32_t> ps, pb;       // padded own / boundary size of the parent's level
  // per (front, boundary slot), level by level with stride b of the level
  std::vector<int32_t> rel;          // index in the parent's front index space (-1 padding)
  std::vector<int32_t> bnd_row;      // padded row of the boundary node (-1 padding)
  // right-hand-side scatter, destination-major: R[rc_dst[q], :] -= sum_s G[rc_src[s], :], s in rc_ptr[q .. q+1)
  std::vector<int32_t> rc_dst, rc_ptr, rc_src;
  std::vector<int64_t> pad_diag;     // A-storage offsets of the diagonal entries of the padding slots (set to 1)
  // pull lists: children of front g = child_ids[child_ptr[g] .. child_ptr[g + 1]); per front (as a child): padded
  // boundary size of its level, offset of its C block, offset of its rel list, and whether it pushes (parent level
  // does not pull)
  std::vector<int32_t> child_ptr, child_ids, f_b;
  std::vector<int64_t> f_offC, f_relOff, f_invOff;
  std::vector<int32_t> inv_rel;      // per child: parent front index -> boundary slot of the child (-1: none)
  std::vector<uint8_t> push;
};

struct HostPlan {
  bool use_tree = false;
  TreePlan tree;
  int64_t N = 0, E = 0, L = 0;
  // compact position space of THIS rank (domain-decomposed plans hold only the rank's own region + the shared top):
  // N0c non-electrode nodes then the nU electrode nodes; Nc = N0c + nU.  parts == 1: N0c = N0, Nc = N.
  int64_t Nc = 0, N0c = 0;
  std::vector<uint8_t> elem_own;     // E: 1 = this rank computes the element's Jacobian columns (all 1 when parts == 1)
  int64_t nU = 0, N0 = 0, nb = 0, max_block = 0, nnz = 0, n_levels = 0;
  // nested dissection level 1: P interiors of <= nI nodes (padded to nI slots each), boundary lists of
  // <= bmax separator nodes; level 2: the nC separator nodes in block-tridiagonal order.
  // Position space: [P*nI interior slots | nC separators | nU electrode nodes] = NP rows.
  int64_t P = 0, nI = 0, bmax = 0, nC = 0, NI = 0, N0P = 0, NP = 0, szPP = 0, szV = 0;
  std::vector<int32_t> bnd_loc;               // P*bmax level-2 local positions (-1 padding)
  std::vector<int64_t> m2_dst;                // destinations (offsets in [Sd|So]-shaped block storage)
  std::vector<int32_t> m2_src_ptr, m2_src;    // sources: p*bmax*bmax + a*bmax + b
  std::vector<int32_t> rc_ptr, rc_src;        // separator local pos -> sources p*bmax + a
  std::vector<int64_t> sg_off;                // P*bmax*bmax offsets into [Sd|So] (-1 padding)
  std::vector<double> nodes;       // N*2
  std::vector<int32_t> elem;       // E*3
  // electrodes
  std::vector<int32_t> el_ptr, el_elems;      // electrode -> elements
  std::vector<int32_t> U;                     // electrode nodes, electrode order, first occurrence
  std::vector<int32_t> u_of;                  // node -> index in U or -1
  std::vector<uint8_t> inB;                   // L*nU: 1 if U-node is Dirichlet for excitation i
  std::vector<int32_t> we_ptr, we_u;          // electrode -> (U index, weight) lists
  std::vector<double> we_w;
  bool u_contiguous = false;                  // electrode m owns U[ue_ptr[m]:ue_ptr[m+1]] exclusively
  std::vector<int32_t> ue_ptr;                // L+1
  std::vector<double> wk;                     // nU: weight of the node inside its electrode
  int32_t bp = 0;                             // max nodes per electrode, rounded up to a multiple of 8
  // ordering
  std::vector<int32_t> pos;                   // node -> COMPACT position (F0 nodes in padded order, then N0+u)
  std::vector<int32_t> ppos;                  // node -> padded position (factor stage)
  std::vector<int32_t> row_map;               // N0: compact F0 row -> padded row
  std::vector<int32_t> node_at;               // position -> node
  std::vector<int32_t> blk_ptr;               // nb+1 (level-2 local positions)
  std::vector<int32_t> blk_of;                // nC: block of level-2 local position
  std::vector<int64_t> offD, offS;            // nb (+1): element offsets of diagonal / sub-diagonal blocks
  int64_t szD = 0, szS = 0;                   // total elements in diag / sub-diag storage
  // CSR pattern of K (original numbering) and deterministic gather lists
  std::vector<int32_t> rowptr, colidx;        // N+1, nnz
  std::vector<int32_t> src_ptr, src;          // nnz+1, 9E: source code = e*9 + a*3 + b
  std::vector<int64_t> slot_dst;              // nnz: offset into the dense destination or -1
  std::vector<uint8_t> slot_kind;             // 0 none, 1 diag block, 2 sub block, 3 K0U, 4 KUU
  // per-element gather offsets into the selected inverse [Sd | So] for pairs
  // (0,0),(1,1),(2,2),(0,1),(1,2),(2,0); -1 => 0
  std::vector<int64_t> g0_off;                // E*6
  std::vector<int32_t> elem_pos;              // E*3 node positions
  // element patches for the Woodbury kernel: elements that share nodes are grouped so that a CTA stages
  // each node row of Yh / Xh in shared memory once (<= patch_node_cap distinct nodes per patch)
  int32_t patch_node_cap = 0;
  std::vector<int32_t> patch_eptr, patch_elems;    // patch -> element ids
  std::vector<int32_t> patch_nptr, patch_nodes;    // patch -> node POSITIONS
  std::vector<uint8_t> patch_local;                // 3 per patch element (patch order): local node index
};

// (Re)build the element patches: <= node_cap distinct nodes (<= 255) and <= elem_cap elements per patch.
void build_patches(HostPlan& p, int32_t node_cap, int32_t elem_cap);

// Throws std::runtime_error on invalid input.
void build_host_plan(HostPlan& p, const double* nodes, const int64_t* elem, int64_t N, int64_t E,
                     const int64_t* el_ptr, const int64_t* el_elems, int64_t L,
                     int64_t target_block, int64_t nd_interior, bool nd_tree = false, int nd_pull = 1, int nd_ways = 4,
                     int parts = 1, int part = 0);

}  // namespace ceit
