// C ABI of libceit_b200 (see include/ceit_b200.h) and the host-side orchestration of the kernels.
#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/ceit_b200.h"
#include "dense.cuh"
#include "factor_tree.h"
#include "gemm.cuh"
#include "kernels.cuh"
#include "plan.h"
#include "scalar.cuh"

namespace ceit {
std::atomic<int64_t> g_launches{0};
int g_force_simt = 0;
int g_gemm_tile = 0;
double g_flops = 0.0;
int g_gemm_vec = 1;
int g_pdl = 1;              // programmatic dependent launch of the GEMM / tile kernels (gemm.cuh: launch_pdl)
int g_gemm_min_ctas64 = 300;   // measured (tools/stage_breakdown.py --opt gemm_min_ctas64=...): 120 -> 300 is 2 % at 50_50 / 200x200
static thread_local std::string g_err;
}  // namespace ceit

using namespace ceit;

#define CEIT_FAIL(code, msg)     \
  do {                           \
    ceit::g_err = (msg);         \
    return (code);               \
  } while (0)
#define CUDA_TRY(expr)                                                                              \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess) {                                                                        \
      ceit::g_err = std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" __FILE__ ":" +        \
                    std::to_string(__LINE__) + ")";                                                 \
      return CEIT_ERR_CUDA;                                                                         \
    }                                                                                               \
  } while (0)
#define KERNEL_CHECK() CUDA_TRY(cudaGetLastError())

namespace {

// ---- optional per-stage device timers (cudaEvent pairs on the caller's stream) ---------------------
// Enabled with ceit_set_option("stage_timers", 1); read (and reset) with ceit_stage_times(), which
// synchronises.  Stages: 0 assemble, 1 factor (base stage), 2 per-excitation stage, 3 Woodbury kernel,
// 4 inverse model, 5 reconstruct, 6 forward extras.
constexpr int N_STAGES = 8;
int g_stage_timers = 0;
int g_no_patch = 0;
int g_no_lowrank = 0;
int g_shard_parts = 1, g_shard_part = 0;   // plans created afterwards: domain decomposition (ranks, this rank)
int g_wb_persist = 1;   // 1: persistent pipelined low-rank Woodbury kernel when the shapes allow it (even nU)
int g_ex_overlap = 1;   // excitation stage: ZW / base potentials on a second stream beside the Yh update
int g_ru_wc = 0, g_ru_cpt = 0, g_ru_waves = 8;   // rank_update_kernel layout overrides (tuning; 0 = auto)
int g_direct_excitation = 0;
int g_patch_nodes = 0;   // experiment: override the node cap of the element patches
int64_t g_nd_interior = 64;   // target nodes per nested-dissection interior / leaf; <= 0 disables the dissection
int g_assemble_gather = 0;    // 1: the slot-gather form of K1 (one thread per CSR slot re-derives the geometry)
int g_nd_tree = 1;            // 1: multi-level nested dissection (factor_tree.h); 0: interiors + block-tridiagonal chain
int g_nd_pull = 1;            // plans created afterwards: 1 = every level pulls its children's update matrices (one
                              // destination-centric launch per level), 0 = children push (one pass per sibling index)
int g_nd_ways = 4;            // plans created afterwards: 4 = two cuts per front (half as many levels), 2 = one cut
int g_front_kernel = 0;       // 1: levels of small fronts run the fused front kernel.  Measured on MESH_50_50: 25-50 us
                              // per level on ONE SM (global-load latency chains + FP64 products) against ~32 us for the
                              // chip-wide launches it replaces, step 0.77 -> 0.82 ms: a tested option, off by default
struct StageEv {
  int stage;
  cudaEvent_t a, b;
};
std::vector<StageEv> g_stage_events;
double g_stage_flops[N_STAGES] = {0, 0, 0, 0, 0, 0, 0, 0};
struct StageScope {
  int stage;
  cudaStream_t st;
  cudaEvent_t a = nullptr, b = nullptr;
  double f0 = 0.0;
  StageScope(int s, cudaStream_t st_) : stage(s), st(st_) {
    if (!g_stage_timers) return;
    f0 = ceit::g_flops;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaEventRecord(a, st);
  }
  ~StageScope() {
    if (!a) return;
    cudaEventRecord(b, st);
    g_stage_events.push_back({stage, a, b});
    g_stage_flops[stage] += ceit::g_flops - f0;
  }
};

inline unsigned cdiv(int64_t a, int64_t b) { return (unsigned)((a + b - 1) / b); }

// Woodbury patch kernel (one group of <= 32 elements per CTA).  Shared memory: per-node constants
// (nU double4), node rows of Yh and Xh (2 * cap * pitch doubles, odd pitch), partial dots
// (8 x 6 x 32), the solved coefficients (6 x 32) and phi0 at the patch nodes.  The node cap is chosen
// so that two CTAs fit an SM when the rows are short enough (staging of one overlaps compute of the other).
constexpr int64_t WB_SMEM_SM = 226 * 1024;
inline int64_t woodbury_pitch(int64_t nU) { return nU | 1; }   // odd: row bases hit all 16 bank pairs
inline size_t woodbury_patch_smem(int64_t nU, int64_t cap) {
  return (size_t)(nU * 32 + 2 * cap * woodbury_pitch(nU) * 8 + WBP_WARPS * 6 * 32 * 8 + 6 * 32 * 8 + cap * 8 + 64);
}
// low-rank kernel: only the Yh rows are staged (plus 2 bp + 3 small values per node)
// (layout in woodbury_lr_kernel: mbarrier, SoA constants, Yh rows at an even pitch, YB/ZW at stride bp+1, ...)
inline int64_t woodbury_lr_cst_len(int64_t nU, int64_t L) { return 3 * ((nU + 1) & ~1LL) + ((L + 1) & ~1LL); }
inline size_t woodbury_lr_smem(int64_t nU, int64_t cap, int64_t bp, int64_t L) {
  return (size_t)(16 + woodbury_lr_cst_len(nU, L) * 8 + cap * woodbury_lr_pitch((int)nU) * 8 + 2 * cap * (bp + 1) * 8 +
                  2 * cap * 8 + 12 * 32 * 8 + 64);
}
// node cap of a patch for the low-rank kernel: 3 CTAs per SM if that leaves >= 20 nodes (~24 elements) per
// patch, else 2 (measured: occupancy beats patch size, 400x400 mesh 30 -> 18 ms)
inline int woodbury_lr_cap(int64_t nU, int64_t bp, int64_t L) {
  for (int per_sm = 3; per_sm >= 1; --per_sm) {
    const int64_t budget = (227 * 1024) / per_sm - 1024 - 256;
    int64_t cap = 40;
    while (cap > 0 && (int64_t)woodbury_lr_smem(nU, cap, bp, L) > budget) --cap;
    if (cap >= (per_sm == 3 ? 20 : 16) || per_sm == 1) return (int)cap;
  }
  return 0;
}
// persistent pipelined kernel: column chunks (electrode ranges starting at an even column) and the node cap that lets
// WBQ_SLOTS slots fit one SM; fewest chunks that leave room for patches of >= 26 nodes (~32 elements), else the
// split with the largest cap.  Returns false when the kernel does not apply (odd nU, no even split, tiny cap).
inline bool wbq_configure(const HostPlan& h, WbqChunks* out, int* cap_out) {
  if ((h.nU & 1) || !h.u_contiguous || h.bp > 64 || h.L < 2) return false;
  const int64_t budget = 227 * 1024 - 1024;
  WbqChunks best{};
  int best_cap = 0;
  for (int n = 1; n <= 4 && n <= h.L; ++n) {
    WbqChunks c{};
    c.n = n;
    c.el[0] = 0;
    bool ok = true;
    for (int q = 1; q < n && ok; ++q) {
      int m = (int)((h.L * q + n / 2) / n);               // nearest electrode boundary with an even first column
      int found = -1;
      for (int d = 0; d < h.L && found < 0; ++d) {
        if (m + d < h.L && m + d > c.el[q - 1] && (h.ue_ptr[m + d] & 1) == 0) found = m + d;
        else if (m - d > c.el[q - 1] && m - d < h.L && (h.ue_ptr[m - d] & 1) == 0) found = m - d;
      }
      if (found < 0) ok = false;
      c.el[q] = found;
    }
    if (!ok) continue;
    c.el[n] = (int)h.L;
    c.wmax = 0;
    for (int q = 0; q < n; ++q) {
      const int w = h.ue_ptr[c.el[q + 1]] - h.ue_ptr[c.el[q]];
      if (w & 1) ok = false;
      c.wmax = std::max(c.wmax, w);
    }
    if (!ok) continue;
    int cap = 40;
    while (cap > 0 && wbq_smem_bytes(c.wmax, (int)h.L, cap, (int)h.bp) > budget) --cap;
    if (cap > best_cap) {
      best_cap = cap;
      best = c;
    }
    if (cap >= 26) break;
  }
  if (best_cap < 16) return false;
  *out = best;
  *cap_out = best_cap;
  return true;
}
inline int64_t zw_ld(int64_t nbatch, int64_t bp) { return (nbatch * (bp + 1) + 1) & ~(int64_t)1; }
inline int woodbury_patch_cap(int64_t nU) {
  for (int per_sm = 2; per_sm >= 1; --per_sm) {
    int64_t budget = WB_SMEM_SM / per_sm - 1024;
    int64_t cap = 40;                       // 32 elements never need more than ~34 distinct nodes
    while (cap > 0 && (int64_t)woodbury_patch_smem(nU, cap) > budget) --cap;
    if (cap >= 20 || per_sm == 1) return (int)cap;
  }
  return 0;
}

}  // namespace

struct ceit_plan {
  HostPlan h;
  void* xg = nullptr;             // decomposed plans: export slots of the sub-tree roots' update matrices (all ranks)
  WbqPatchRec* d_patch_rec = nullptr;   // per (patch, lane): element id, packed local node slots, 2 omega c A factor
  int32_t* d_exp_b = nullptr;
  int64_t* d_exp_offC = nullptr;
  bool wbq_ok = false;            // the persistent pipelined Woodbury kernel applies (patches were built for it)
  WbqChunks wbq{};
  int sm_count = 0;
  bool uploaded = false;
  size_t dev_bytes = 0;
  std::vector<void*> allocs;
  // static tables
  double* d_nodes = nullptr;
  int32_t* d_elem = nullptr;
  int32_t* d_src_ptr = nullptr;
  int32_t* d_src = nullptr;
  int32_t* d_elem_dst = nullptr;   // inverse of d_src: (element, a, b) -> position in the slot-sorted contribution list
  void* d_contrib = nullptr;       // 9E (re, im) pairs
  int64_t* d_slot_dst = nullptr;
  uint8_t* d_slot_kind = nullptr;
  int32_t* d_pos = nullptr;
  int32_t* d_elem_pos = nullptr;
  int64_t* d_g0_off = nullptr;
  uint8_t* d_inB = nullptr;
  int32_t* d_we_ptr = nullptr;
  int32_t* d_we_u = nullptr;
  double* d_we_w = nullptr;
  double* d_area = nullptr;
  int32_t *d_patch_eptr = nullptr, *d_patch_elems = nullptr, *d_patch_nptr = nullptr, *d_patch_nodes = nullptr;
  uint8_t* d_patch_local = nullptr;
  int n_patches = 0;
  int32_t* d_ue_ptr = nullptr;
  double* d_wk = nullptr;
  cd* d_vals = nullptr;  // CSR values of the last assemble
  int* d_info = nullptr;
  bool assembled = false;
  // numeric state (type-erased; element size esz = 8 real / 16 complex)
  int mode = 0;  // 0 none, 1 real, 2 complex
  void *Ad = nullptr, *Bs = nullptr, *Ld = nullptr, *Linv = nullptr, *Lsub = nullptr, *Sel = nullptr;
  void *K0U = nullptr, *Zw = nullptr, *Xh = nullptr, *XP = nullptr, *KUU = nullptr, *SU = nullptr, *Ptmp = nullptr;
  int32_t* d_row_map = nullptr;
  void *tmpPanel = nullptr, *g0e = nullptr, *tmpT = nullptr, *tmpPanel2 = nullptr, *Ptmp2 = nullptr;
  void *Aint = nullptr, *Lint = nullptr, *LintInv = nullptr, *Ainv = nullptr, *Bt = nullptr, *Wn = nullptr;
  void *Tm = nullptr, *G1 = nullptr;
  void *tAct = nullptr, *YP = nullptr;   // Y_T through a second backward sweep: t of the active rows, padded Y_T
  void *Mreg = nullptr, *LT = nullptr, *LinvT = nullptr, *Tinv = nullptr, *te = nullptr, *YT = nullptr, *ye = nullptr;
  double* d_gamma = nullptr;
  void* splitk_ws = nullptr;
  // auxiliary stream + events: independent chains of a stage (right-hand-side sweep vs selected inverse, ...)
  // run concurrently and join before the stage ends; also valid under stream capture (fork/join in the graph)
  cudaStream_t s2 = nullptr, s3 = nullptr, s4 = nullptr, s5 = nullptr;
  SideLane lane_a, lane_b, lane_c;   // Linv row blocks of the blocked factorisations on st / on s3 / on s4
  cudaEvent_t ev_ring[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  int ev_next = 0;
  void* g1e = nullptr;
  void* ZW = nullptr;
  void *TBB = nullptr, *LB = nullptr, *LinvB = nullptr, *Dcat = nullptr;
  int32_t *d_node_at = nullptr, *d_bnd_loc = nullptr, *d_m2_src_ptr = nullptr, *d_m2_src = nullptr;
  int32_t *d_rc_ptr = nullptr, *d_rc_src = nullptr;
  int64_t *d_m2_dst = nullptr, *d_sg_off = nullptr;
  // multi-level nested dissection tables (TreePlan)
  int32_t *d_t_sib = nullptr, *d_t_ps = nullptr, *d_t_pb = nullptr, *d_t_rel = nullptr, *d_t_bnd_row = nullptr;
  int32_t *d_t_rc_dst = nullptr, *d_t_rc_ptr = nullptr, *d_t_rc_src = nullptr;
  int64_t *d_t_pA = nullptr, *d_t_pB = nullptr, *d_t_pC = nullptr, *d_t_pad = nullptr;
  int32_t *d_t_child_ptr = nullptr, *d_t_child_ids = nullptr, *d_t_fb = nullptr, *d_t_inv_rel = nullptr;
  int64_t *d_t_foffC = nullptr, *d_t_frelOff = nullptr, *d_t_finvOff = nullptr;
  uint8_t* d_t_push = nullptr;
  int alloc_mode = 0;
  int64_t generation = 0;   // bumped whenever a numeric / excitation workspace is (re)allocated: graphs captured
                            // before hold stale pointers (ceit_plan_generation)
  // per-excitation workspace
  int ex_cap = 0, ex_mode = 0;
  void *St = nullptr, *Ls = nullptr, *Linvs = nullptr, *Sinv = nullptr, *Yh = nullptr, *phi0 = nullptr;
  void *tvec = nullptr, *tmpS = nullptr;
  double* d_elec = nullptr;
  double* d_node_abs = nullptr;
};

namespace {

int dev_alloc(ceit_plan* p, void** out, size_t bytes) {
  if (bytes == 0) bytes = 16;
  CUDA_TRY(cudaMalloc(out, bytes));
  p->allocs.push_back(*out);
  p->dev_bytes += bytes;
  return 0;
}

template <class V>
int upload_vec(ceit_plan* p, V** dptr, const std::vector<V>& v) {
  int rc = dev_alloc(p, (void**)dptr, v.size() * sizeof(V));
  if (rc) return rc;
  if (!v.empty()) CUDA_TRY(cudaMemcpy(*dptr, v.data(), v.size() * sizeof(V), cudaMemcpyHostToDevice));
  return 0;
}

void dev_free_tracked(ceit_plan* p, void*& ptr) {
  if (!ptr) return;
  for (auto& a : p->allocs)
    if (a == ptr) {
      cudaFree(a);
      a = nullptr;
    }
  ptr = nullptr;
}

int upload_plan(ceit_plan* p) {
  HostPlan& h = p->h;
  int rc;
  if ((rc = upload_vec(p, &p->d_nodes, h.nodes))) return rc;
  if ((rc = upload_vec(p, &p->d_elem, h.elem))) return rc;
  if ((rc = upload_vec(p, &p->d_src_ptr, h.src_ptr))) return rc;
  if ((rc = upload_vec(p, &p->d_src, h.src))) return rc;
  {
    std::vector<int32_t> inv(h.src.size());
    for (size_t q = 0; q < h.src.size(); ++q) inv[(size_t)h.src[q]] = (int32_t)q;
    if ((rc = upload_vec(p, &p->d_elem_dst, inv))) return rc;
    if ((rc = dev_alloc(p, &p->d_contrib, h.src.size() * 2 * sizeof(double)))) return rc;
  }
  if ((rc = upload_vec(p, &p->d_exp_b, h.tree.exp_b))) return rc;
  if ((rc = upload_vec(p, &p->d_exp_offC, h.tree.exp_offC))) return rc;
  if ((rc = upload_vec(p, &p->d_slot_dst, h.slot_dst))) return rc;
  if ((rc = upload_vec(p, &p->d_slot_kind, h.slot_kind))) return rc;
  if ((rc = upload_vec(p, &p->d_pos, h.pos))) return rc;
  if ((rc = upload_vec(p, &p->d_elem_pos, h.elem_pos))) return rc;
  if ((rc = upload_vec(p, &p->d_g0_off, h.g0_off))) return rc;
  if ((rc = upload_vec(p, &p->d_inB, h.inB))) return rc;
  if ((rc = upload_vec(p, &p->d_we_ptr, h.we_ptr))) return rc;
  if ((rc = upload_vec(p, &p->d_we_u, h.we_u))) return rc;
  if ((rc = upload_vec(p, &p->d_we_w, h.we_w))) return rc;
  if ((rc = upload_vec(p, &p->d_patch_eptr, h.patch_eptr))) return rc;
  if ((rc = upload_vec(p, &p->d_patch_elems, h.patch_elems))) return rc;
  if ((rc = upload_vec(p, &p->d_patch_nptr, h.patch_nptr))) return rc;
  if ((rc = upload_vec(p, &p->d_patch_nodes, h.patch_nodes))) return rc;
  if ((rc = upload_vec(p, &p->d_patch_local, h.patch_local))) return rc;
  p->n_patches = (int)h.patch_eptr.size() - 1;
  if ((rc = upload_vec(p, &p->d_node_at, h.node_at))) return rc;
  if ((rc = upload_vec(p, &p->d_row_map, h.row_map))) return rc;
  if ((rc = upload_vec(p, &p->d_bnd_loc, h.bnd_loc))) return rc;
  if ((rc = upload_vec(p, &p->d_m2_dst, h.m2_dst))) return rc;
  if ((rc = upload_vec(p, &p->d_m2_src_ptr, h.m2_src_ptr))) return rc;
  if ((rc = upload_vec(p, &p->d_m2_src, h.m2_src))) return rc;
  if ((rc = upload_vec(p, &p->d_rc_ptr, h.rc_ptr))) return rc;
  if ((rc = upload_vec(p, &p->d_rc_src, h.rc_src))) return rc;
  if ((rc = upload_vec(p, &p->d_sg_off, h.sg_off))) return rc;
  if (h.use_tree) {
    const TreePlan& t = h.tree;
    if ((rc = upload_vec(p, &p->d_t_sib, t.sib))) return rc;
    if ((rc = upload_vec(p, &p->d_t_ps, t.ps))) return rc;
    if ((rc = upload_vec(p, &p->d_t_pb, t.pb))) return rc;
    if ((rc = upload_vec(p, &p->d_t_rel, t.rel))) return rc;
    if ((rc = upload_vec(p, &p->d_t_bnd_row, t.bnd_row))) return rc;
    if ((rc = upload_vec(p, &p->d_t_rc_dst, t.rc_dst))) return rc;
    if ((rc = upload_vec(p, &p->d_t_rc_ptr, t.rc_ptr))) return rc;
    if ((rc = upload_vec(p, &p->d_t_rc_src, t.rc_src))) return rc;
    if ((rc = upload_vec(p, &p->d_t_pA, t.pA))) return rc;
    if ((rc = upload_vec(p, &p->d_t_pB, t.pB))) return rc;
    if ((rc = upload_vec(p, &p->d_t_pC, t.pC))) return rc;
    if ((rc = upload_vec(p, &p->d_t_pad, t.pad_diag))) return rc;
    if ((rc = upload_vec(p, &p->d_t_child_ptr, t.child_ptr))) return rc;
    if ((rc = upload_vec(p, &p->d_t_child_ids, t.child_ids))) return rc;
    if ((rc = upload_vec(p, &p->d_t_fb, t.f_b))) return rc;
    if ((rc = upload_vec(p, &p->d_t_foffC, t.f_offC))) return rc;
    if ((rc = upload_vec(p, &p->d_t_frelOff, t.f_relOff))) return rc;
    if ((rc = upload_vec(p, &p->d_t_push, t.push))) return rc;
    if ((rc = upload_vec(p, &p->d_t_inv_rel, t.inv_rel))) return rc;
    if ((rc = upload_vec(p, &p->d_t_finvOff, t.f_invOff))) return rc;
  }
  if ((rc = upload_vec(p, &p->d_ue_ptr, h.ue_ptr))) return rc;
  if ((rc = upload_vec(p, &p->d_wk, h.wk))) return rc;
  if ((rc = dev_alloc(p, (void**)&p->d_area, h.E * sizeof(double)))) return rc;
  if ((rc = dev_alloc(p, (void**)&p->d_vals, h.nnz * sizeof(cd)))) return rc;
  if ((rc = dev_alloc(p, (void**)&p->d_info, sizeof(int)))) return rc;
  CUDA_TRY(cudaMemset(p->d_info, 0, sizeof(int)));
  elem_param_kernel<<<cdiv(h.E, 256), 256>>>((int)h.E, p->d_nodes, p->d_elem, nullptr, p->d_area);
  CEIT_COUNT_LAUNCH();
  if (p->n_patches > 0) {
    if ((rc = dev_alloc(p, (void**)&p->d_patch_rec, (size_t)p->n_patches * 32 * sizeof(WbqPatchRec)))) return rc;
    patch_rec_kernel<<<cdiv((int64_t)p->n_patches * 32, 256), 256>>>(p->n_patches, p->d_patch_eptr, p->d_patch_elems,
                                                                     p->d_patch_local, p->d_area, p->d_patch_rec);
    CEIT_COUNT_LAUNCH();
  }
  KERNEL_CHECK();
  CUDA_TRY(cudaDeviceSynchronize());
  p->uploaded = true;
  return 0;
}

inline bool rank_update_shape_ok(const HostPlan& h) {
  return h.u_contiguous && h.bp <= 64 && h.nU <= 1024 && h.nU > h.bp;
}
inline bool use_rank_update(const ceit_plan* p) { return !g_direct_excitation && rank_update_shape_ok(p->h); }
// Y_T = Xh T as a second backward sweep over the elimination forest instead of an N x nU x nU product: pays when the
// product is large (400x400: 9.5e10 flops = 3.2 ms against ~1 ms of sweep); "yt_sweep" = 0 / 1 forces it off / on
int g_yt_sweep = -1;
inline bool yt_by_sweep(const ceit_plan* p) {
  if (!p->h.use_tree || !use_rank_update(p)) return false;
  if (g_yt_sweep >= 0) return g_yt_sweep != 0;
  return 2.0 * (double)p->h.Nc * (double)p->h.nU * (double)p->h.nU > 2.0e10;
}

// workspaces are never (re)allocated while the caller's stream is being captured: cudaFree / cudaMalloc are not
// capturable and a replay would not redo them.  The caller runs the sequence eagerly once first.
inline int refuse_alloc_in_capture(cudaStream_t st, const char* what) {
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cs) == cudaSuccess && cs != cudaStreamCaptureStatusNone) {
    ceit::g_err = std::string(what) + ": workspace allocation needed during stream capture; run the same call "
                  "sequence eagerly once before capturing it";
    return CEIT_ERR_STATE;
  }
  return 0;
}

template <class T>
int alloc_numeric(ceit_plan* p, cudaStream_t st) {
  const int want = Sc<T>::is_complex ? 2 : 1;
  if (p->alloc_mode == want) return 0;
  if (int rc_ = refuse_alloc_in_capture(st, "ceit_factor")) return rc_;
  CUDA_TRY(cudaDeviceSynchronize());   // buffers of the other mode may still be in use on any stream
  p->generation++;
  void** bufs[] = {&p->xg, &p->Ad, &p->Bs, &p->Ld, &p->Linv, &p->Lsub, &p->Sel, &p->K0U, &p->Zw,
                   &p->Xh, &p->KUU, &p->SU, &p->Ptmp, &p->tmpPanel, &p->tmpT, &p->tmpPanel2, &p->Ptmp2, &p->g0e, &p->Aint, &p->Lint,
                   &p->LintInv, &p->Ainv, &p->Bt, &p->Wn, &p->Tm, &p->G1, &p->XP, &p->Mreg, &p->LT,
                   &p->LinvT, &p->Tinv, &p->te, &p->YT, &p->ye, &p->splitk_ws, &p->g1e, &p->tAct, &p->YP};
  for (auto b : bufs) dev_free_tracked(p, *b);
  // excitation workspace depends on the type too
  void** ex[] = {&p->St, &p->Ls, &p->Linvs, &p->Sinv, &p->Yh, &p->phi0, &p->tvec, &p->tmpS,
                 &p->TBB, &p->LB, &p->LinvB, &p->Dcat, &p->ZW};
  for (auto b : ex) dev_free_tracked(p, *b);
  p->ex_cap = 0;
  p->alloc_mode = 0;
  const HostPlan& h = p->h;
  const size_t es = sizeof(T);
  int64_t maxS = 0;
  for (int64_t k = 0; k + 1 < h.nb; ++k) maxS = std::max<int64_t>(maxS, h.offS[k + 1] - h.offS[k]);
  int rc;
  if ((rc = dev_alloc(p, &p->Ad, h.szD * es))) return rc;
  if ((rc = dev_alloc(p, &p->Bs, h.szS * es))) return rc;
  if ((rc = dev_alloc(p, &p->Ld, h.szD * es))) return rc;
  if ((rc = dev_alloc(p, &p->Linv, h.szD * es))) return rc;
  if ((rc = dev_alloc(p, &p->Lsub, h.szS * es))) return rc;
  if ((rc = dev_alloc(p, &p->Sel, (h.szD + h.szS + h.szPP + h.szV) * es))) return rc;
  if ((rc = dev_alloc(p, &p->K0U, (h.N0P + h.nU) * h.nU * es))) return rc;   // + nU rows: partial S_U of a decomposed plan
  if ((rc = dev_alloc(p, &p->xg, (int64_t)h.tree.parts * h.tree.roots_per_rank * h.tree.slot_sz * es))) return rc;
  if ((rc = dev_alloc(p, &p->Zw, (h.use_tree ? 0 : h.N0P * h.nU) * es))) return rc;
  if ((rc = dev_alloc(p, &p->Xh, h.Nc * h.nU * es))) return rc;
  if ((rc = dev_alloc(p, &p->XP, h.N0P * h.nU * es))) return rc;
  if ((rc = dev_alloc(p, &p->Aint, h.szPP * es))) return rc;
  if ((rc = dev_alloc(p, &p->Lint, h.szPP * es))) return rc;
  if ((rc = dev_alloc(p, &p->LintInv, h.szPP * es))) return rc;
  if ((rc = dev_alloc(p, &p->Ainv, h.szPP * es))) return rc;
  if ((rc = dev_alloc(p, &p->Bt, h.szV * es))) return rc;
  if ((rc = dev_alloc(p, &p->Wn, h.szV * es))) return rc;
  if ((rc = dev_alloc(p, &p->Tm, std::max<int64_t>(h.P * h.bmax * h.bmax, h.tree.szC) * es))) return rc;
  if ((rc = dev_alloc(p, &p->G1, std::max<int64_t>(h.P * h.bmax, h.tree.maxG) * h.nU * es))) return rc;
  if ((rc = dev_alloc(p, &p->Mreg, h.nU * h.nU * es))) return rc;
  if ((rc = dev_alloc(p, &p->LT, h.nU * h.nU * es))) return rc;
  if ((rc = dev_alloc(p, &p->LinvT, h.nU * h.nU * es))) return rc;
  if ((rc = dev_alloc(p, &p->Tinv, h.nU * h.nU * es))) return rc;
  if ((rc = dev_alloc(p, &p->te, h.nU * es))) return rc;
  if ((rc = dev_alloc(p, &p->YT, (h.Nc * h.nU + 2) * es))) return rc;
  if ((rc = dev_alloc(p, &p->ye, h.Nc * es))) return rc;
  if (!p->d_gamma && (rc = dev_alloc(p, (void**)&p->d_gamma, sizeof(double)))) return rc;
  if ((rc = dev_alloc(p, &p->g1e, 6 * h.E * es))) return rc;
  if (yt_by_sweep(p)) {
    if ((rc = dev_alloc(p, &p->tAct, h.tree.rowsA * h.nU * es))) return rc;
    if ((rc = dev_alloc(p, &p->YP, h.N0P * h.nU * es))) return rc;
  }
  if ((rc = dev_alloc(p, &p->splitk_ws, (size_t)splitk_max_slices(h.nU, h.nU, h.N0P) * h.nU * h.nU * sizeof(double)))) return rc;
  if ((rc = dev_alloc(p, &p->KUU, h.nU * h.nU * es))) return rc;
  if ((rc = dev_alloc(p, &p->SU, h.nU * h.nU * es))) return rc;
  if ((rc = dev_alloc(p, &p->Ptmp, maxS * es))) return rc;
  if ((rc = dev_alloc(p, &p->Ptmp2, maxS * es))) return rc;
  int64_t panel = (int64_t)TILE * h.max_block;
  for (const TreeLevel& lv : h.tree.lv)
    if (lv.s > TILE) panel = std::max<int64_t>(panel, (int64_t)lv.n * TILE * lv.s);
  if ((rc = dev_alloc(p, &p->tmpPanel, panel * es))) return rc;
  if ((rc = dev_alloc(p, &p->tmpT, (int64_t)TILE * h.nU * es))) return rc;
  if ((rc = dev_alloc(p, &p->tmpPanel2, (int64_t)TILE * h.max_block * es))) return rc;
  if ((rc = dev_alloc(p, &p->g0e, 6 * h.E * es))) return rc;
  p->alloc_mode = want;
  return 0;
}

template <class T>
int alloc_excitation_ws(ceit_plan* p, int want_cap, cudaStream_t st) {
  const int want_mode = Sc<T>::is_complex ? 2 : 1;
  if (p->ex_cap >= want_cap && p->ex_mode == want_mode) return 0;
  if (int rc_ = refuse_alloc_in_capture(st, "ceit_jacobian_block / ceit_forward")) return rc_;
  CUDA_TRY(cudaDeviceSynchronize());
  p->generation++;
  void** ex[] = {&p->St, &p->Ls, &p->Linvs, &p->Sinv, &p->Yh, &p->phi0, &p->tvec, &p->tmpS,
                 &p->TBB, &p->LB, &p->LinvB, &p->Dcat, &p->ZW};
  for (auto b : ex) dev_free_tracked(p, *b);
  if (p->d_elec) { void* q = p->d_elec; dev_free_tracked(p, q); p->d_elec = nullptr; }
  const HostPlan& h = p->h;
  const size_t es = sizeof(T);
  const int64_t c = want_cap, u2 = h.nU * h.nU;
  int rc;
  if ((rc = dev_alloc(p, &p->St, c * u2 * es))) return rc;
  if ((rc = dev_alloc(p, &p->Ls, c * u2 * es))) return rc;
  if ((rc = dev_alloc(p, &p->Linvs, c * u2 * es))) return rc;
  if ((rc = dev_alloc(p, &p->Sinv, c * u2 * es))) return rc;
  if ((rc = dev_alloc(p, &p->Yh, c * h.Nc * h.nU * es))) return rc;
  if ((rc = dev_alloc(p, &p->phi0, c * h.Nc * es))) return rc;
  if ((rc = dev_alloc(p, &p->tvec, c * h.nU * es))) return rc;
  if ((rc = dev_alloc(p, &p->tmpS, c * (int64_t)TILE * h.nU * es))) return rc;
  if ((rc = dev_alloc(p, &p->TBB, c * (int64_t)h.bp * h.bp * es))) return rc;
  if ((rc = dev_alloc(p, &p->LB, c * (int64_t)h.bp * h.bp * es))) return rc;
  if ((rc = dev_alloc(p, &p->LinvB, c * (int64_t)h.bp * h.bp * es))) return rc;
  if ((rc = dev_alloc(p, &p->Dcat, c * (int64_t)(h.bp + 1) * h.nU * es))) return rc;
  if ((rc = dev_alloc(p, &p->ZW, (h.Nc * zw_ld(c, h.bp) + 2) * es))) return rc;   // + 2: aligned windows may read one past
  if ((rc = dev_alloc(p, (void**)&p->d_elec, c * h.L * sizeof(double)))) return rc;
  p->ex_cap = want_cap;
  p->ex_mode = want_mode;
  return 0;
}

// the rank-(b+1) excitation path needs contiguous electrode ranges in U, <= 64 nodes per electrode and
// nU <= 1024; otherwise every excitation factorises its own padded S_U (excitation_direct)
int g_multi_stream = 1;
// `to` waits for everything enqueued on `from` so far
inline int stream_after(ceit_plan* p, cudaStream_t from, cudaStream_t to) {
  if (from == to) return 0;
  cudaEvent_t e = p->ev_ring[p->ev_next];
  p->ev_next = (p->ev_next + 1) & 7;
  CUDA_TRY(cudaEventRecord(e, from));
  CUDA_TRY(cudaStreamWaitEvent(to, e, 0));
  return 0;
}
inline int ensure_aux_stream(ceit_plan* p) {
  if (p->s2) return 0;
  CUDA_TRY(cudaStreamCreateWithFlags(&p->s2, cudaStreamNonBlocking));
  CUDA_TRY(cudaStreamCreateWithFlags(&p->s3, cudaStreamNonBlocking));
  CUDA_TRY(cudaStreamCreateWithFlags(&p->s4, cudaStreamNonBlocking));
  CUDA_TRY(cudaStreamCreateWithFlags(&p->s5, cudaStreamNonBlocking));
  CUDA_TRY(p->lane_c.init());
  CUDA_TRY(p->lane_a.init());
  CUDA_TRY(p->lane_b.init());
  for (auto& e : p->ev_ring) CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  return 0;
}
#define CEIT_AFTER(from, to) do { int rc_ = stream_after(p, from, to); if (rc_) return rc_; } while (0)

// lane for the plan-less inverse-model entry points (one per calling thread)
inline SideLane* inv_lane() {
  static thread_local SideLane lane;
  if (!g_multi_stream) return nullptr;
  return lane.init() == cudaSuccess ? &lane : nullptr;
}

// excitation-independent part of the element blocks for the low-rank Woodbury kernel (real path only)
inline void launch_g1(ceit_plan* p, const double* g0e, const double* YT, const double* Xh, cudaStream_t st) {
  const HostPlan& h = p->h;
  pdl(g1_kernel, cdiv(h.E, 8), 256, 0, st)(h.E, (int)h.nU, p->d_elem_pos, g0e, YT, Xh, (double*)p->g1e);
  CEIT_COUNT_LAUNCH();
}
inline void launch_g1(ceit_plan*, const cd*, const cd*, const cd*, cudaStream_t) {}

inline void schur_u_gemm(ceit_plan* p, cudaStream_t st, int64_t nU, int64_t N0P, const double* K0U, const double* X,
                         double* SU) {
  // K_0U^T K00^-1 K_0U is symmetric: only the lower triangle is accumulated; symmetrize_su() mirrors it afterwards
  gemm_splitk(st, 1, 0, nU, nU, N0P, -1.0, K0U, nU, X, nU, 1.0, SU, nU, (double*)p->splitk_ws, /*lower_only=*/1);
}
inline void symmetrize_su(cudaStream_t st, int64_t nU, double* SU) {
  pdl(symmetrize_from_lower_kernel, cdiv(nU * nU, 256), 256, 0, st)(nU, SU, nU);
  CEIT_COUNT_LAUNCH();
}
inline void symmetrize_su(cudaStream_t, int64_t, cd*) {}   // the complex path forms the full product
inline void schur_u_gemm(ceit_plan*, cudaStream_t st, int64_t nU, int64_t N0P, const cd* K0U, const cd* X, cd* SU) {
  gemm<cd>(st, 1, 0, nU, nU, N0P, neg(Sc<cd>::one()), K0U, nU, 0, X, nU, 0, Sc<cd>::one(), SU, nU, 0);
}

// T = (S_U + gamma e e^T)^-1 and t_e = T e on stream s_t: shared by every excitation (rank-update excitation stage)
template <class T>
int base_shared_inverse(ceit_plan* p, cudaStream_t s_t) {
  const HostPlan& h = p->h;
  const int64_t nU = h.nU;
  const size_t es = sizeof(T);
  const T one = Sc<T>::one(), zero = Sc<T>::zero();
  T* SU = (T*)p->SU;
  if (use_rank_update(p)) {
    // T, t_e = T e: shared by every excitation (Y_T = Xh T and y_e = Y_T e follow once X is complete)
    T *Mreg = (T*)p->Mreg, *LT = (T*)p->LT, *LinvT = (T*)p->LinvT, *Tinv = (T*)p->Tinv, *te = (T*)p->te;
    pdl(trace_gamma_kernel<T>, 1, 256, 0, s_t)((int)nU, SU, p->d_gamma);
    CEIT_COUNT_LAUNCH();
    pdl(add_rank1_e_kernel<T>, cdiv(nU * nU, 256), 256, 0, s_t)((int)nU, SU, p->d_gamma, Mreg);
    CEIT_COUNT_LAUNCH();
    // (LT and LinvT were zeroed at the start of the stage, off this chain)
    potrf_trtri_blocked<T>(s_t, nU, Mreg, nU, 0, LT, nU, 0, LinvT, nU, 0, (T*)p->tmpT, 0, 1, p->d_info,
                           g_multi_stream ? &p->lane_b : nullptr);
    gemm<T>(s_t, 1, 0, nU, nU, nU, one, LinvT, nU, 0, LinvT, nU, 0, zero, Tinv, nU, 0);
    pdl(rowsum_e_kernel<T>, cdiv(nU, 8), 256, 0, s_t)(nU, (int)nU, Tinv, te);
    CEIT_COUNT_LAUNCH();
  }
  return 0;
}

// end of the base stage: element blocks of the selected inverse, compact X, Xh = [X ; -I], Y_T = Xh T, G1
template <class T>
int base_tail(ceit_plan* p, cudaStream_t st, cudaStream_t sr, cudaStream_t s_t) {
  const HostPlan& h = p->h;
  const int64_t nU = h.nU;
  const T one = Sc<T>::one(), zero = Sc<T>::zero();
  T *Sel = (T*)p->Sel, *g0e = (T*)p->g0e, *Xh = (T*)p->XP, *Xc = (T*)p->Xh;
  pdl(gather_g0_kernel<T>, cdiv(6 * h.E, 256), 256, 0, st)(6 * h.E, p->d_g0_off, Sel, g0e);
  CEIT_COUNT_LAUNCH();
  pdl(compact_rows_kernel<T>, cdiv(h.N0c, 8), 256, 0, sr)(h.N0c, (int)nU, p->d_row_map, Xh, Xc);
  CEIT_COUNT_LAUNCH();
  pdl(fill_neg_identity_kernel<T>, cdiv(nU * nU, 256), 256, 0, sr)((int)nU, Xc + h.N0c * nU);
  CEIT_COUNT_LAUNCH();
  if (use_rank_update(p)) {
    T *Tinv = (T*)p->Tinv, *YT = (T*)p->YT, *ye = (T*)p->ye;
    CEIT_AFTER(s_t, sr);
    if (yt_by_sweep(p)) {
      // the padded Y_T rows come from the second backward sweep (factor_impl); rows of U: Xh[U] T = -T
      pdl(compact_rows_kernel<T>, cdiv(h.N0c, 8), 256, 0, sr)(h.N0c, (int)nU, p->d_row_map, (const T*)p->YP, YT);
      CEIT_COUNT_LAUNCH();
      pdl(negate_copy_kernel<T>, cdiv(nU * nU, 256), 256, 0, sr)(nU * nU, Tinv, YT + h.N0c * nU);
      CEIT_COUNT_LAUNCH();
    } else {
      gemm<T>(sr, 0, 0, h.Nc, nU, nU, one, Xc, nU, 0, Tinv, nU, 0, zero, YT, nU, 0);               // Y_T = Xh T
    }
    pdl(rowsum_e_kernel<T>, cdiv(h.Nc, 8), 256, 0, sr)(h.Nc, (int)nU, YT, ye);
    CEIT_COUNT_LAUNCH();
    CEIT_AFTER(sr, st);
    launch_g1(p, g0e, YT, Xc, st);
  } else {
    CEIT_AFTER(s_t, sr);
    CEIT_AFTER(sr, st);
  }
  return 0;
}

// ---- CUDA backend of factor_tree.h: batched DMMA GEMMs, the tile factorisation and the gather / scatter kernels on
// two streams (lane 0 = the caller's stream: matrix chain; lane 1: right-hand-side chain) --------------------------
template <class T>
struct CudaTreeBK {
  ceit_plan* p;
  cudaStream_t lane[3];
  T* panel;
  int rc = 0;
  bool lane2_used = false;
  void after(int from, int to) {
    if (from == 2 && !lane2_used) return;      // nothing was forked to lane 2 (every level ran the fused front step)
    if (to == 2) lane2_used = true;
    if (!rc) rc = stream_after(p, lane[from], lane[to]);
  }
  FrontFusedArgs front_args(const TreeLevel& lv) const {
    return FrontFusedArgs{lv.s, lv.b, lv.f0, p->d_t_child_ptr, p->d_t_child_ids, p->d_t_fb, p->d_t_rel, p->d_t_foffC,
                          p->d_t_frelOff, p->d_t_inv_rel, p->d_t_finvOff};
  }
  bool can_fuse() const { return !Sc<T>::is_complex && g_front_kernel; }
  void front_fused(int ln, const TreePlan&, const TreeLevel& lv, T* A, T* Bt, T* C, T* Ainv, T* W) {
    launch_front_fused(ln, lv, A, Bt, C, Ainv, W);
  }
  void launch_front_fused(int ln, const TreeLevel& lv, double* A, double* Bt, double* C, double* Ainv, double* W) {
    static bool attr_done = false;
    if (!attr_done) {
      cudaFuncSetAttribute(front_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FRONT_SMEM);
      attr_done = true;
    }
    pdl(front_fused_kernel, lv.n, 256, FRONT_SMEM, lane[ln])(front_args(lv), A + lv.offA, Bt + lv.offB, C, C + lv.offC,
                                                            Ainv + lv.offA, W + lv.offB, p->d_info);
    CEIT_COUNT_LAUNCH();
    g_flops += (double)lv.n * ((2.0 / 3.0 + 1.0) * lv.s * lv.s * lv.s + 3.0 * lv.s * lv.s * lv.b +
                               (double)lv.s * lv.b * lv.b);
  }
  void launch_front_fused(int, const TreeLevel&, cd*, cd*, cd*, cd*, cd*) {}   // complex: pull + separate steps
  void pull_children(int ln, const TreePlan&, const TreeLevel& lv, T* A, T* Bt, T* C) {
    if (lv.f0 == 0 && p->h.tree.child_ptr[lv.n] == 0) return;       // the leaf level has no children
    const int64_t span = (int64_t)lv.s + lv.b;
    dim3 grid(cdiv(span * span, 256), lv.n);
    pdl(pull_children_kernel<T>, grid, 256, 0, lane[ln])(front_args(lv), A + lv.offA, Bt + lv.offB, C, C + lv.offC);
    CEIT_COUNT_LAUNCH();
  }
  void gemm(int ln, int opA, int opB, int64_t M, int64_t N, int64_t K, T alpha, const T* A, int64_t lda, int64_t sA,
            const T* B, int64_t ldb, int64_t sB, T beta, T* C, int64_t ldc, int64_t sC, int batch, int lower_only) {
    ceit::gemm<T>(lane[ln], opA, opB, M, N, K, alpha, A, lda, sA, B, ldb, sB, beta, C, ldc, sC, batch, lower_only);
  }
  void potrf(int ln, int64_t n, T* A, T* L, T* Li, int batch) {
    potrf_trtri_blocked<T>(lane[ln], n, A, n, n * n, L, n, n * n, Li, n, n * n, panel, (int64_t)TILE * n, batch,
                           p->d_info, (g_multi_stream && ln == 0) ? &p->lane_a : nullptr);
  }
  void extend_add(int ln, const TreePlan&, const TreeLevel& lv, int pass, const T* U, T* A, T* Bt, T* C) {
    dim3 grid(cdiv((int64_t)lv.b * lv.b, 256), lv.n);
    pdl(extend_add_kernel<T>, grid, 256, 0, lane[ln])(lv.b, pass, p->d_t_push + lv.f0, p->d_t_sib + lv.f0, p->d_t_pA + lv.f0,
                                                     p->d_t_pB + lv.f0, p->d_t_pC + lv.f0, p->d_t_ps + lv.f0,
                                                     p->d_t_pb + lv.f0, p->d_t_rel + lv.offR, U, A, Bt, C);
    CEIT_COUNT_LAUNCH();
  }
  void extend_gather(int ln, const TreePlan&, const TreeLevel& lv, const T* Sg, const T* V, T* C) {
    dim3 grid(cdiv((int64_t)lv.b * lv.b, 256), lv.n);
    pdl(extend_gather_kernel<T>, grid, 256, 0, lane[ln])(lv.b, p->d_t_pA + lv.f0, p->d_t_pB + lv.f0, p->d_t_pC + lv.f0,
                                                        p->d_t_ps + lv.f0, p->d_t_pb + lv.f0, p->d_t_rel + lv.offR, Sg,
                                                        V, C, C + lv.offC);
    CEIT_COUNT_LAUNCH();
  }
  void rhs_scatter(int ln, const TreePlan&, const TreeLevel& lv, int64_t nU, const T* G, T* R) {
    if (lv.rc_n == 0) return;
    pdl(rhs_scatter_kernel<T>, dim3((unsigned)lv.rc_n, cdiv(nU, 128)), 128, 0, lane[ln])(
        lv.rc_n, (int)nU, p->d_t_rc_dst + lv.rc0, p->d_t_rc_ptr + lv.rc0, p->d_t_rc_src, G, R);
    CEIT_COUNT_LAUNCH();
  }
  void zero_rows(int ln, T* ptr, int64_t count) { cudaMemsetAsync(ptr, 0, (size_t)count * sizeof(T), lane[ln]); }
  void rows_gather(int ln, const TreePlan&, const TreeLevel& lv, int64_t nU, const T* Xp, T* G) {
    const int64_t rows = (int64_t)lv.n * lv.b;
    pdl(bnd_rows_gather_kernel<T>, cdiv(rows, 8), 256, 0, lane[ln])(rows, (int)nU, p->d_t_bnd_row + lv.offR, Xp, G);
    CEIT_COUNT_LAUNCH();
  }
};

template <class T>
TreeBufs<T> tree_bufs(ceit_plan* p) {
  const TreePlan& t = p->h.tree;
  TreeBufs<T> m;
  m.A = (T*)p->Aint; m.Lf = (T*)p->Lint; m.Linv = (T*)p->LintInv; m.Ainv = (T*)p->Ainv; m.Bt = (T*)p->Bt; m.W = (T*)p->Wn;
  m.SigPP = (T*)p->Sel; m.V = (T*)p->Sel + t.szA; m.C = (T*)p->Tm; m.R = (T*)p->K0U; m.Xp = (T*)p->XP; m.G = (T*)p->G1;
  return m;
}

// partial S_U of a decomposed plan: P = -K_0U[rows]^T t[rows] (lower triangle; the slab was zeroed with K0U)
inline void schur_u_partial(ceit_plan* p, cudaStream_t st, int64_t nU, int64_t rows, const double* K0U, const double* X,
                            double* P) {
  gemm_splitk(st, 1, 0, nU, nU, rows, -1.0, K0U, nU, X, nU, 0.0, P, nU, (double*)p->splitk_ws, /*lower_only=*/1);
}
inline void schur_u_partial(ceit_plan*, cudaStream_t, int64_t, int64_t, const cd*, const cd*, cd*) {}

// common tail of the tree base stage once S_U (lower triangle) is complete on `sr`: T on its own stream, the
// top-down passes, Y_T, the element blocks
template <class T>
int factor_tree_finish(ceit_plan* p, cudaStream_t st, cudaStream_t sr, CudaTreeBK<T>& bk) {
  const HostPlan& h = p->h;
  const TreePlan& t = h.tree;
  const int64_t nU = h.nU;
  const size_t es = sizeof(T);
  const T one = Sc<T>::one(), zero = Sc<T>::zero(), mone = neg(Sc<T>::one());
  T *Sel = (T*)p->Sel, *Ainv = (T*)p->Ainv, *Xh = (T*)p->XP, *Wn = (T*)p->Wn, *G1 = (T*)p->G1, *SU = (T*)p->SU;
  TreeBufs<T> m = tree_bufs<T>(p);
  int rc;
  symmetrize_su(sr, nU, SU);
  cudaStream_t s_t = sr;
  if (g_multi_stream) {
    s_t = p->s3;
    CEIT_AFTER(sr, s_t);
  }
  if ((rc = base_shared_inverse<T>(p, s_t))) return rc;
  CEIT_AFTER(sr, st);                                   // W of every level, A^-1
  CUDA_TRY(cudaMemcpyAsync(Sel, Ainv, t.szA * es, cudaMemcpyDeviceToDevice, st));
  const bool yt_sweep = yt_by_sweep(p);
  // active rows this rank holds: everything, or (decomposed) its own block and the top's
  const int64_t own0 = t.parts > 1 ? t.ra0[t.part] : 0, own1 = t.parts > 1 ? t.ra1[t.part] : t.rowsA;
  const int64_t top0 = t.parts > 1 ? t.ra_top0 : t.rowsA;
  if (yt_sweep) {                                       // the X sweep overwrites t in place: keep the active rows
    T* tAct = (T*)p->tAct;
    if (own1 > own0)
      CUDA_TRY(cudaMemcpyAsync(tAct + own0 * nU, Xh + own0 * nU, (own1 - own0) * nU * es, cudaMemcpyDeviceToDevice, sr));
    if (t.rowsA > top0)
      CUDA_TRY(cudaMemcpyAsync(tAct + top0 * nU, Xh + top0 * nU, (t.rowsA - top0) * nU * es, cudaMemcpyDeviceToDevice, sr));
  }
  tree_top_down<T>(t, nU, m, bk, one, zero, mone);
  if (bk.rc) return bk.rc;
  if (yt_sweep) {
    // Y_T = Xh T without the N x nU x nU product: (t T) over the active rows, then the same backward sweep
    T *YP = (T*)p->YP, *Tinv = (T*)p->Tinv, *tAct = (T*)p->tAct;
    CEIT_AFTER(s_t, sr);
    if (own1 > own0)
      gemm<T>(sr, 0, 0, own1 - own0, nU, nU, one, tAct + own0 * nU, nU, 0, Tinv, nU, 0, zero, YP + own0 * nU, nU, 0);
    if (t.rowsA > top0)
      gemm<T>(sr, 0, 0, t.rowsA - top0, nU, nU, one, tAct + top0 * nU, nU, 0, Tinv, nU, 0, zero, YP + top0 * nU, nU, 0);
    tree_back_sweep<T>(t, nU, Wn, G1, YP, bk, 1, zero, one, mone);
    if (bk.rc) return bk.rc;
  }
  if ((rc = base_tail<T>(p, st, sr, s_t))) return rc;
  KERNEL_CHECK();
  p->mode = Sc<T>::is_complex ? 2 : 1;
  return 0;
}

// second half of a decomposed base stage (after the caller's exchange): import the foreign sub-tree roots' update
// matrices, the top bottom-up, S_U = K_UU + sum of the partial products - the top's own rows, then the common tail
template <class T>
int factor_tree_second_half(ceit_plan* p, cudaStream_t st) {
  const HostPlan& h = p->h;
  const TreePlan& t = h.tree;
  if (!h.use_tree || t.parts <= 1 || p->mode != -1)
    CEIT_FAIL(CEIT_ERR_STATE, "ceit_factor_phase(2): call ceit_factor_phase(1) and the exchange first");
  const int64_t nU = h.nU;
  const T one = Sc<T>::one(), zero = Sc<T>::zero(), mone = neg(Sc<T>::one());
  T *K0U = (T*)p->K0U, *Xh = (T*)p->XP, *SU = (T*)p->SU;
  int rc;
  cudaStream_t sr = st;
  if (g_multi_stream) {
    if ((rc = ensure_aux_stream(p))) return rc;
    sr = p->s2;
  }
  TreeBufs<T> m = tree_bufs<T>(p);
  CudaTreeBK<T> bk{p, {st, sr, g_multi_stream ? p->s4 : sr}, (T*)p->tmpPanel};
  if (t.roots_per_rank > 0) {
    dim3 grid(cdiv(t.slot_sz, 256), t.parts * t.roots_per_rank);
    pdl(export_copy_kernel<T>, grid, 256, 0, st)(0, t.part * t.roots_per_rank, (t.part + 1) * t.roots_per_rank, p->d_exp_b,
                                                p->d_exp_offC, t.slot_sz, (T*)p->Tm, (T*)p->xg, 0);
    CEIT_COUNT_LAUNCH();
  }
  bk.after(0, 1);                                       // the exchange results and the imports reach every lane
  bk.after(0, 2);
  tree_bottom_up<T>(t, nU, m, bk, one, zero, mone, t.n_own_segs, t.H);
  bk.after(2, 1);
  if (bk.rc) return bk.rc;
  pdl(add_lower_kernel<T>, cdiv(nU * nU, 256), 256, 0, sr)(nU * nU, K0U + t.rowsA * nU, SU);
  CEIT_COUNT_LAUNCH();
  if (t.rowsA > t.ra_top0)
    schur_u_gemm(p, sr, nU, t.rowsA - t.ra_top0, K0U + t.ra_top0 * nU, Xh + t.ra_top0 * nU, SU);
  return factor_tree_finish<T>(p, st, sr, bk);
}

// ---- base stage -----------------------------------------------------------------------------------
// K00 = K[F0,F0] in the ordering [interiors | separators]:
//   level 1: the P mutually disconnected interiors (<= 64 nodes each) are factorised and eliminated as
//            ONE batch (one tile kernel + a handful of batched GEMMs);
//   level 2: the separator Schur complement M2 = K_CC - sum_p Bt_p^T A_p^-1 Bt_p is block-tridiagonal
//            over its BFS level sets: block Cholesky, nU right-hand sides, selected inverse;
//   then the interior parts of X and of the selected inverse follow from batched GEMMs.
// phase: 0 = the whole stage (plans without domain decomposition); decomposed plans (TreePlan::parts > 1) run it in two
// halves around the caller's NCCL exchange (ceit_factor_phase): 1 = assembly scatter, the rank's own segments
// bottom-up, its partial S_U, export of the sub-tree roots' update matrices; 2 = import, the top, S_U, T, top-down.
template <class T>
int factor_impl(ceit_plan* p, cudaStream_t st, int phase = 0) {
  const HostPlan& h = p->h;
  int rc = alloc_numeric<T>(p, st);
  if (rc) return rc;
  if (phase == 2) return factor_tree_second_half<T>(p, st);
  T *Ad = (T*)p->Ad, *Bs = (T*)p->Bs, *Ld = (T*)p->Ld, *Linv = (T*)p->Linv, *Lsub = (T*)p->Lsub;
  T *Sel = (T*)p->Sel, *K0U = (T*)p->K0U, *Zw = (T*)p->Zw, *Xh = (T*)p->XP, *Xc = (T*)p->Xh, *KUU = (T*)p->KUU;   // Xh: padded rows here
  T *SU = (T*)p->SU, *Ptmp = (T*)p->Ptmp, *tmpPanel = (T*)p->tmpPanel, *g0e = (T*)p->g0e;
  T *Aint = (T*)p->Aint, *Lint = (T*)p->Lint, *LintInv = (T*)p->LintInv, *Ainv = (T*)p->Ainv;
  T *Bt = (T*)p->Bt, *Wn = (T*)p->Wn, *Tm = (T*)p->Tm, *G1 = (T*)p->G1;
  T* Sd = Sel;
  T* So = Sel + h.szD;
  T* SigPP = Sel + h.szD + h.szS;
  T* Vneg = SigPP + h.szPP;
  const size_t es = sizeof(T);
  const int64_t nU = h.nU, N0P = h.N0P, NI = h.NI, P = h.P, nI = h.nI, bmax = h.bmax, nC = h.nC;
  const int64_t sPP = nI * nI, sV = nI * bmax, sT = bmax * bmax, sG = bmax * nU, sX = nI * nU;
  const T one = Sc<T>::one(), zero = Sc<T>::zero(), mone = neg(Sc<T>::one());
  CUDA_TRY(cudaMemsetAsync(Ad, 0, h.szD * es, st));
  CUDA_TRY(cudaMemsetAsync(Bs, 0, h.szS * es, st));
  CUDA_TRY(cudaMemsetAsync(Ld, 0, h.szD * es, st));
  CUDA_TRY(cudaMemsetAsync(Linv, 0, h.szD * es, st));
  CUDA_TRY(cudaMemsetAsync(K0U, 0, (h.use_tree ? h.tree.rowsA : N0P) * nU * es, st));
  CUDA_TRY(cudaMemsetAsync(KUU, 0, nU * nU * es, st));
  CUDA_TRY(cudaMemsetAsync(p->d_info, 0, sizeof(int), st));
  if (use_rank_update(p)) {      // factors of the shared inverse T: zero-initialised here, off the critical chain
    CUDA_TRY(cudaMemsetAsync(p->LT, 0, nU * nU * es, st));
    CUDA_TRY(cudaMemsetAsync(p->LinvT, 0, nU * nU * es, st));
  }
  if (P > 0 || h.use_tree) {
    CUDA_TRY(cudaMemsetAsync(Aint, 0, h.szPP * es, st));
    CUDA_TRY(cudaMemsetAsync(Lint, 0, h.szPP * es, st));
    CUDA_TRY(cudaMemsetAsync(LintInv, 0, h.szPP * es, st));
    CUDA_TRY(cudaMemsetAsync(Bt, 0, h.szV * es, st));
    if (h.use_tree) {
      CUDA_TRY(cudaMemsetAsync(Tm, 0, h.tree.szC * es, st));
      const int64_t npad = (int64_t)h.tree.pad_diag.size();
      if (npad > 0) {
        pdl(pad_diag_list_kernel<T>, cdiv(npad, 256), 256, 0, st)(npad, p->d_t_pad, Aint);
        CEIT_COUNT_LAUNCH();
      }
    } else {
      pdl(pad_identity_kernel<T>, cdiv(P * nI, 256), 256, 0, st)((int)P, (int)nI, p->d_node_at, Aint);
      CEIT_COUNT_LAUNCH();
    }
  }
  pdl(scatter_dense_kernel<T>, cdiv(h.nnz, 256), 256, 0, st)((int)h.nnz, p->d_vals, p->d_slot_dst, p->d_slot_kind, Ad,
                                                            Bs, K0U, KUU, Aint, Bt);
  CEIT_COUNT_LAUNCH();
  KERNEL_CHECK();
  CUDA_TRY(cudaMemcpyAsync(SU, KUU, nU * nU * es, cudaMemcpyDeviceToDevice, st));
  // Two chains share the factorisation: the MATRIX chain (factor blocks, selected inverse, G0) stays on `st`,
  // the RIGHT-HAND-SIDE chain (forward / backward sweeps with K_0U, S_U, T, Y_T) runs on `sr`; they only meet
  // at the factor blocks (sr waits for st) and at the very end.  Buffers are disjoint: sr writes Xh, Zw, G1,
  // SU, Xc, Mreg/LT/LinvT/Tinv/te/YT/ye; st writes Ad, Bs, Ld, Linv, Lsub, Sel, Ptmp, Tm, g0e.
  cudaStream_t sr = st;
  if (g_multi_stream) {
    if ((rc = ensure_aux_stream(p))) return rc;
    sr = p->s2;
  }
  if (h.use_tree) {
    // ---- multi-level nested dissection (factor_tree.h): bottom-up, S_U -> T on its own stream, top-down ----------
    const TreePlan& t = h.tree;
    TreeBufs<T> m = tree_bufs<T>(p);
    CudaTreeBK<T> bk{p, {st, sr, g_multi_stream ? p->s4 : sr}, tmpPanel};
    if (phase == 1) {
      // the assembled right-hand-side rows of the top count once in the all-reduce: rank 0 keeps them
      if (t.part != 0 && t.rowsA > t.ra_top0)
        CUDA_TRY(cudaMemsetAsync(K0U + t.ra_top0 * nU, 0, (t.rowsA - t.ra_top0) * nU * es, st));
      // the partial-S_U slab behind the active rows: a rank without active fronts (an interior region: no electrode
      // nearby) contributes zeros, and the product below writes the lower triangle only
      CUDA_TRY(cudaMemsetAsync(K0U + t.rowsA * nU, 0, nU * nU * es, st));
      tree_bottom_up<T>(t, nU, m, bk, one, zero, mone, 0, t.n_own_segs);
      if (bk.rc) return bk.rc;
      const int64_t r0 = t.ra0[t.part], r1 = t.ra1[t.part];
      if (r1 > r0)                        // partial S_U over this rank's active rows -> the nU rows after the active block
        schur_u_partial(p, sr, nU, r1 - r0, K0U + r0 * nU, Xh + r0 * nU, K0U + t.rowsA * nU);
      if (t.roots_per_rank > 0) {
        dim3 grid(cdiv(t.slot_sz, 256), t.roots_per_rank);
        pdl(export_copy_kernel<T>, grid, 256, 0, st)(t.part * t.roots_per_rank, -1, -1, p->d_exp_b, p->d_exp_offC, t.slot_sz,
                                                    (T*)p->Tm, (T*)p->xg, 1);
        CEIT_COUNT_LAUNCH();
      }
      bk.after(2, 1);
      CEIT_AFTER(sr, st);
      KERNEL_CHECK();
      p->mode = -1;                        // half-factorised: only ceit_factor_phase(2) may follow
      return 0;
    }
    tree_bottom_up<T>(t, nU, m, bk, one, zero, mone);
    bk.after(2, 1);                                       // W of every level: needed by both top-down chains
    if (bk.rc) return bk.rc;
    KERNEL_CHECK();
    schur_u_gemm(p, sr, nU, t.rowsA, K0U, Xh, SU);      // S_U = K_UU - r^T t over the active rows of all levels
    if ((rc = factor_tree_finish<T>(p, st, sr, bk))) return rc;
    return 0;
  }
  // ---- level 1: eliminate the interiors (batched) -----------------------------------------------------
  if (P > 0) {
    potrf_trtri_blocked<T>(st, nI, Aint, nI, sPP, Lint, nI, sPP, LintInv, nI, sPP, nullptr, 0, (int)P, p->d_info);
    gemm<T>(st, 1, 0, nI, nI, nI, one, LintInv, nI, sPP, LintInv, nI, sPP, zero, Ainv, nI, sPP, (int)P);   // A_p^-1
    CEIT_AFTER(st, sr);
    gemm<T>(st, 0, 0, nI, bmax, nI, one, Ainv, nI, sPP, Bt, bmax, sV, zero, Wn, bmax, sV, (int)P);          // W_p
    gemm<T>(st, 1, 0, bmax, bmax, nI, one, Bt, bmax, sV, Wn, bmax, sV, zero, Tm, bmax, sT, (int)P);         // Bt^T W
    const int64_t nd = (int64_t)h.m2_dst.size();
    if (nd > 0) {
      pdl(m2_gather_kernel<T>, cdiv(nd, 256), 256, 0, st)(nd, p->d_m2_dst, p->d_m2_src_ptr, p->d_m2_src, Tm, h.szD, Ad, Bs);
      CEIT_COUNT_LAUNCH();
    }
    // right-hand sides: t_I = A^-1 K_IU (kept in the interior rows of Xh), r_C = K_CU - Bt^T t_I
    gemm<T>(sr, 0, 0, nI, nU, nI, one, Ainv, nI, sPP, K0U, nU, sX, zero, Xh, nU, sX, (int)P);
    gemm<T>(sr, 1, 0, bmax, nU, nI, one, Bt, bmax, sV, Xh, nU, sX, zero, G1, nU, sG, (int)P);
    pdl(rc_gather_kernel<T>, cdiv(nC, 8), 256, 0, sr)(nC, (int)nU, p->d_rc_ptr, p->d_rc_src, G1, K0U + NI * nU,
                                                     Zw + NI * nU);
    CEIT_COUNT_LAUNCH();
  } else {
    CEIT_AFTER(st, sr);
    CUDA_TRY(cudaMemcpyAsync(Zw, K0U, N0P * nU * es, cudaMemcpyDeviceToDevice, sr));
  }
  KERNEL_CHECK();
  // ---- level 2: block-tridiagonal Cholesky of M2 + forward substitution ----------------------------
  // The chain of nb blocks is eliminated from BOTH ends towards the middle block m ("burn at both ends"): blocks
  // 0..m-1 downwards (Lsub_k = B_k Linv_k^T, as in a plain block Cholesky) on st/sr, blocks nb-1..m+1 upwards
  // (Vsub_k = Linv_k B_{k-1}, D_{k-1} -= Vsub_k^T Vsub_k) on sb/srb, so the serial spine is nb/2 + 1 blocks long.
  auto n_of = [&](int64_t k) { return (int64_t)(h.blk_ptr[k + 1] - h.blk_ptr[k]); };
  auto row_of = [&](int64_t k) { return NI + (int64_t)h.blk_ptr[k]; };
  const int64_t nbk = h.nb, mid = nbk / 2;
  cudaStream_t sb = st, srb = sr;
  if (g_multi_stream && nbk >= 3) {
    sb = p->s4;
    srb = p->s5;
    CEIT_AFTER(st, sb);
    CEIT_AFTER(sr, srb);
  }
  auto factor_block = [&](int64_t k, cudaStream_t sm, cudaStream_t sx, T* tmp, SideLane* lane) -> int {
    const int64_t nk = n_of(k);
    potrf_trtri_blocked<T>(sm, nk, Ad + h.offD[k], nk, 0, Ld + h.offD[k], nk, 0, Linv + h.offD[k], nk, 0, tmp, 0, 1,
                           p->d_info, g_multi_stream ? lane : nullptr);
    CEIT_AFTER(sm, sx);                  // Linv_k (and the coupling factors before it, W_p) are ready for the sweep
    return 0;
  };
  for (int64_t k = 0; k < mid; ++k) {                                  // downward arm
    const int64_t nk = n_of(k), r0 = row_of(k), nn = n_of(k + 1);
    T* Lik = Linv + h.offD[k];
    if ((rc = factor_block(k, st, sr, tmpPanel, &p->lane_a))) return rc;
    if (k > 0) {
      const int64_t np = n_of(k - 1), rp = row_of(k - 1);
      // RHS_k -= Lsub_{k-1} Z_{k-1}   (Z_{k-1} lives in Xh rows of block k-1 during the forward sweep)
      gemm<T>(sr, 0, 0, nk, nU, np, mone, Lsub + h.offS[k - 1], np, 0, Xh + rp * nU, nU, 0, one, Zw + r0 * nU, nU, 0);
    }
    gemm<T>(sr, 0, 0, nk, nU, nk, one, Lik, nk, 0, Zw + r0 * nU, nU, 0, zero, Xh + r0 * nU, nU, 0);   // Z_k = Linv_k RHS_k
    // Lsub_k = B_k Linv_k^T ; A_{k+1} -= Lsub_k Lsub_k^T
    gemm<T>(st, 0, 1, nn, nk, nk, one, Bs + h.offS[k], nk, 0, Lik, nk, 0, zero, Lsub + h.offS[k], nk, 0);
    gemm<T>(st, 0, 1, nn, nn, nk, mone, Lsub + h.offS[k], nk, 0, Lsub + h.offS[k], nk, 0, one, Ad + h.offD[k + 1], nn, 0,
            1, /*lower_only=*/1);
  }
  for (int64_t k = nbk - 1; k > mid; --k) {                            // upward arm
    const int64_t nk = n_of(k), r0 = row_of(k), np = n_of(k - 1);
    T* Lik = Linv + h.offD[k];
    if ((rc = factor_block(k, sb, srb, (T*)p->tmpPanel2, &p->lane_c))) return rc;
    if (k + 1 < nbk) {
      const int64_t nn = n_of(k + 1), rn = row_of(k + 1);
      // RHS_k -= Vsub_{k+1}^T Z_{k+1}
      gemm<T>(srb, 1, 0, nk, nU, nn, mone, Lsub + h.offS[k], nk, 0, Xh + rn * nU, nU, 0, one, Zw + r0 * nU, nU, 0);
    }
    gemm<T>(srb, 0, 0, nk, nU, nk, one, Lik, nk, 0, Zw + r0 * nU, nU, 0, zero, Xh + r0 * nU, nU, 0);
    // Vsub_k = Linv_k B_{k-1} (n_k x n_{k-1}, stored in the slot of B_{k-1}) ; A_{k-1} -= Vsub_k^T Vsub_k
    gemm<T>(sb, 0, 0, nk, np, nk, one, Lik, nk, 0, Bs + h.offS[k - 1], np, 0, zero, Lsub + h.offS[k - 1], np, 0);
    // The update of the MIDDLE block is issued on st after the arms have joined: the downward arm accumulates into
    // the same block (a read-modify-write from two streams loses updates).
    if (k - 1 != mid)
      gemm<T>(sb, 1, 0, np, np, nk, mone, Lsub + h.offS[k - 1], np, 0, Lsub + h.offS[k - 1], np, 0, one,
              Ad + h.offD[k - 1], np, 0, 1, /*lower_only=*/1);
  }
  if (nbk > 0) {                                                       // middle block: both arms meet
    const int64_t nk = n_of(mid), r0 = row_of(mid);
    T* Lik = Linv + h.offD[mid];
    CEIT_AFTER(sb, st);
    CEIT_AFTER(srb, sr);
    if (mid + 1 < nbk) {
      const int64_t nn = n_of(mid + 1);
      gemm<T>(st, 1, 0, nk, nk, nn, mone, Lsub + h.offS[mid], nk, 0, Lsub + h.offS[mid], nk, 0, one, Ad + h.offD[mid], nk,
              0, 1, /*lower_only=*/1);
    }
    if ((rc = factor_block(mid, st, sr, tmpPanel, &p->lane_a))) return rc;
    if (mid > 0) {
      const int64_t np = n_of(mid - 1), rp = row_of(mid - 1);
      gemm<T>(sr, 0, 0, nk, nU, np, mone, Lsub + h.offS[mid - 1], np, 0, Xh + rp * nU, nU, 0, one, Zw + r0 * nU, nU, 0);
    }
    if (mid + 1 < nbk) {
      const int64_t nn = n_of(mid + 1), rn = row_of(mid + 1);
      gemm<T>(sr, 1, 0, nk, nU, nn, mone, Lsub + h.offS[mid], nk, 0, Xh + rn * nU, nU, 0, one, Zw + r0 * nU, nU, 0);
    }
    gemm<T>(sr, 0, 0, nk, nU, nk, one, Lik, nk, 0, Zw + r0 * nU, nU, 0, zero, Xh + r0 * nU, nU, 0);
  }
  // S_U = K_UU - K_0U^T K00^-1 K_0U = K_UU - K_IU^T t_I - Z_C^T Z_C is complete after the FORWARD sweep (t_I and Z
  // still sit in Xh), so T = (S_U + gamma e e^T)^-1, the longest chain of the stage, starts here on its own
  // stream (s_t) beside the backward sweep (sr) and the selected inverse (st).  Small outputs with long inner
  // dimensions: split-K on the real path.
  if (P > 0) schur_u_gemm(p, sr, nU, NI, K0U, Xh, SU);
  if (nC > 0) schur_u_gemm(p, sr, nU, nC, Xh + NI * nU, Xh + NI * nU, SU);
  if (P > 0 || nC > 0) symmetrize_su(sr, nU, SU);
  cudaStream_t s_t = sr;
  if (g_multi_stream) {
    s_t = p->s3;
    CEIT_AFTER(sr, s_t);
  }
  if ((rc = base_shared_inverse<T>(p, s_t))) return rc;
  // backward sweep and selected inverse, from the middle block outwards along both arms:
  //   X_k = Linv_k^T (Z_k - Lsub_k^T X_{k+1}) (k < m),  X_k = Linv_k^T (Z_k - Vsub_k X_{k-1}) (k > m)
  //   k < m: P = Lsub_k Linv_k,  So_k = -Sd_{k+1} P,      Sd_k = Linv_k^T Linv_k - P^T So_k
  //   k > m: R = Linv_k^T Vsub_k, So_{k-1} = -R Sd_{k-1}, Sd_k = Linv_k^T Linv_k - So_{k-1} R^T
  auto back_block = [&](int64_t k, cudaStream_t sx) -> int {       // X_k from Z_k (after the coupling term was applied)
    const int64_t nk = n_of(k), r0 = row_of(k);
    // in-place is not allowed in the GEMM: go through Zw rows of this block
    CUDA_TRY(cudaMemcpyAsync(Zw + r0 * nU, Xh + r0 * nU, nk * nU * es, cudaMemcpyDeviceToDevice, sx));
    gemm<T>(sx, 1, 0, nk, nU, nk, one, Linv + h.offD[k], nk, 0, Zw + r0 * nU, nU, 0, zero, Xh + r0 * nU, nU, 0);
    return 0;
  };
  if (nbk > 0) {
    if ((rc = back_block(mid, sr))) return rc;
    gemm<T>(st, 1, 0, n_of(mid), n_of(mid), n_of(mid), one, Linv + h.offD[mid], n_of(mid), 0, Linv + h.offD[mid], n_of(mid),
            0, zero, Sd + h.offD[mid], n_of(mid), 0);
    CEIT_AFTER(sr, srb);
    CEIT_AFTER(st, sb);
  }
  for (int64_t k = mid - 1; k >= 0; --k) {                             // downward arm, outwards
    const int64_t nk = n_of(k), r0 = row_of(k), nn = n_of(k + 1), rn = row_of(k + 1);
    T* Lik = Linv + h.offD[k];
    gemm<T>(sr, 1, 0, nk, nU, nn, mone, Lsub + h.offS[k], nk, 0, Xh + rn * nU, nU, 0, one, Xh + r0 * nU, nU, 0);
    if ((rc = back_block(k, sr))) return rc;
    gemm<T>(st, 1, 0, nk, nk, nk, one, Lik, nk, 0, Lik, nk, 0, zero, Sd + h.offD[k], nk, 0);
    gemm<T>(st, 0, 0, nn, nk, nk, one, Lsub + h.offS[k], nk, 0, Lik, nk, 0, zero, Ptmp, nk, 0);
    gemm<T>(st, 0, 0, nn, nk, nn, mone, Sd + h.offD[k + 1], nn, 0, Ptmp, nk, 0, zero, So + h.offS[k], nk, 0);
    gemm<T>(st, 1, 0, nk, nk, nn, mone, Ptmp, nk, 0, So + h.offS[k], nk, 0, one, Sd + h.offD[k], nk, 0);
  }
  for (int64_t k = mid + 1; k < nbk; ++k) {                            // upward arm, outwards
    const int64_t nk = n_of(k), r0 = row_of(k), np = n_of(k - 1), rp = row_of(k - 1);
    T* Lik = Linv + h.offD[k];
    T* Rt = (T*)p->Ptmp2;
    gemm<T>(srb, 0, 0, nk, nU, np, mone, Lsub + h.offS[k - 1], np, 0, Xh + rp * nU, nU, 0, one, Xh + r0 * nU, nU, 0);
    if ((rc = back_block(k, srb))) return rc;
    gemm<T>(sb, 1, 0, nk, nk, nk, one, Lik, nk, 0, Lik, nk, 0, zero, Sd + h.offD[k], nk, 0);
    gemm<T>(sb, 1, 0, nk, np, nk, one, Lik, nk, 0, Lsub + h.offS[k - 1], np, 0, zero, Rt, np, 0);
    gemm<T>(sb, 0, 0, nk, np, np, mone, Rt, np, 0, Sd + h.offD[k - 1], np, 0, zero, So + h.offS[k - 1], np, 0);
    gemm<T>(sb, 0, 1, nk, nk, np, mone, So + h.offS[k - 1], np, 0, Rt, np, 0, one, Sd + h.offD[k], nk, 0);
  }
  CEIT_AFTER(srb, sr);
  CEIT_AFTER(sb, st);
  // ---- back to the interiors: x_I = t_I - W x_C[bnd] (sr), selected inverse blocks (st) ----------------
  if (P > 0) {
    if (h.nb == 0) CEIT_AFTER(st, sr);   // W_p (no separator block ordered the streams yet)
    pdl(bnd_rows_gather_kernel<T>, cdiv(P * bmax, 8), 256, 0, sr)(P * bmax, (int)nU, p->d_bnd_loc, Xh + NI * nU, G1);
    CEIT_COUNT_LAUNCH();
    gemm<T>(sr, 0, 0, nI, nU, bmax, mone, Wn, bmax, sV, G1, nU, sG, one, Xh, nU, sX, (int)P);
    pdl(gather_g0_kernel<T>, cdiv(P * sT, 256), 256, 0, st)(P * sT, p->d_sg_off, Sel, Tm);        // Sigma2 on the cliques
    CEIT_COUNT_LAUNCH();
    gemm<T>(st, 0, 0, nI, bmax, bmax, mone, Wn, bmax, sV, Tm, bmax, sT, zero, Vneg, bmax, sV, (int)P);   // -W Sg
    CUDA_TRY(cudaMemcpyAsync(SigPP, Ainv, h.szPP * es, cudaMemcpyDeviceToDevice, st));
    gemm<T>(st, 0, 1, nI, nI, bmax, mone, Vneg, bmax, sV, Wn, bmax, sV, one, SigPP, nI, sPP, (int)P);    // + V W^T
  }
  if ((rc = base_tail<T>(p, st, sr, s_t))) return rc;
  KERNEL_CHECK();
  p->mode = Sc<T>::is_complex ? 2 : 1;
  return 0;
}

// ---- per-excitation stage for excitations [i0, i0+nb) into workspace slots [0, nb) -------------------
template <class T>
int excitation_impl(ceit_plan* p, int i0, int nbatch, cudaStream_t st) {
  const HostPlan& h = p->h;
  const int64_t nU = h.nU, N = h.Nc, N0 = h.N0c, u2 = nU * nU;   // compact position space
  T *St = (T*)p->St, *Ls = (T*)p->Ls, *Linvs = (T*)p->Linvs, *Sinv = (T*)p->Sinv, *Yh = (T*)p->Yh;
  T *phi0 = (T*)p->phi0, *tvec = (T*)p->tvec, *tmpS = (T*)p->tmpS, *Xh = (T*)p->Xh, *SU = (T*)p->SU;
  const T one = Sc<T>::one(), zero = Sc<T>::zero();
  const size_t es = sizeof(T);
  cudaStream_t sx = st;
  if (use_rank_update(p)) {
    // rank-(b+1) update of the shared inverse T (kernels.cuh): no per-excitation nU x nU factorisation
    const int bp = h.bp;
    T *TBB = (T*)p->TBB, *LB = (T*)p->LB, *LinvB = (T*)p->LinvB, *Dcat = (T*)p->Dcat;
    T *Tinv = (T*)p->Tinv, *te = (T*)p->te, *YT = (T*)p->YT, *ye = (T*)p->ye;
    pdl(gather_tbb_kernel<T>, dim3(cdiv(bp * bp, 256), nbatch), 256, 0, st)((int)nU, bp, Tinv, p->d_ue_ptr, i0, TBB);
    CEIT_COUNT_LAUNCH();
    CUDA_TRY(cudaMemsetAsync(LB, 0, (int64_t)nbatch * bp * bp * es, st));
    CUDA_TRY(cudaMemsetAsync(LinvB, 0, (int64_t)nbatch * bp * bp * es, st));
    potrf_trtri_blocked<T>(st, bp, TBB, bp, (int64_t)bp * bp, LB, bp, (int64_t)bp * bp, LinvB, bp, (int64_t)bp * bp,
                           nullptr, 0, nbatch, p->d_info);
    const size_t sm1_base = (size_t)(2 * bp * bp + bp + nU + 256 + 8 * bp) * es;
    const int stage_rows = sm1_base + (size_t)bp * nU * es <= 200 * 1024;     // T[B_i,:] in shared memory if it fits
    const size_t sm1 = sm1_base + (stage_rows ? (size_t)bp * nU * es : 0);
    static bool attr_done = false;
    if (!attr_done) {
      CUDA_TRY(cudaFuncSetAttribute(excite_small_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
#define CEIT_RU_ATTR(c, w) \
  CUDA_TRY(cudaFuncSetAttribute(rank_update_kernel<T, c, w>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024))
      CEIT_RU_ATTR(1, 1); CEIT_RU_ATTR(2, 1); CEIT_RU_ATTR(3, 1);
      CEIT_RU_ATTR(1, 2); CEIT_RU_ATTR(2, 2); CEIT_RU_ATTR(3, 2);
      CEIT_RU_ATTR(1, 4); CEIT_RU_ATTR(2, 4); CEIT_RU_ATTR(3, 4);
#undef CEIT_RU_ATTR
      attr_done = true;
    }
    pdl(excite_small_kernel<T>, nbatch, 256, sm1, st)((int)nU, bp, Tinv, te, LinvB, p->d_ue_ptr, i0, p->d_gamma, Dcat,
                                                    stage_rows);
    CEIT_COUNT_LAUNCH();
    auto update = [&](cudaStream_t su, int64_t rows, const T* src, const T* yv, int zero_rows, T* out, int64_t stride) {
      // column chunks of 32 WC CPT (WC warps across, CPT columns per thread): least cost over all chunks, where a
      // padded column costs max(shared-memory wavefronts, FP64 issue) per useful FMA of a thread's RH x CPT tile
      // (CPT = 3: 14/24, 2: 12/16, 1: 10/8 - the RH broadcast loads of the A column are shared by CPT columns);
      // ties go to the widest chunk (D_i is staged once per CTA, fewer re-reads of the A columns)
      int best_wc = 4, best_cpt = 3, best_nz = (int)cdiv(nU, 384);
      double best_cost = 1e300;
      for (int wc : {4, 2, 1})
        for (int cpt = 3; cpt >= 1; --cpt) {
          const int w = 32 * wc * cpt;
          const int nzc = (int)cdiv(nU, w);
          const double cost = (double)nzc * w * (cpt == 3 ? 14.0 / 24 : (cpt == 2 ? 12.0 / 16 : 10.0 / 8));
          if (cost < best_cost * (1 - 1e-9)) { best_cost = cost; best_wc = wc; best_cpt = cpt; best_nz = nzc; }
        }
      if (g_ru_wc > 0 && g_ru_cpt > 0) {
        best_wc = g_ru_wc; best_cpt = g_ru_cpt; best_nz = (int)cdiv(nU, 32 * best_wc * best_cpt);
      }
      const int wc = best_wc, cpt = best_cpt, nz = best_nz;
      const int cw = (int)std::min<int64_t>(32 * wc * cpt, nU);              // chunk width (last chunk may be short)
      const int rt = (Sc<T>::is_complex ? 4 : 8) * (8 / wc);                  // rows per pass
      const size_t smu = (size_t)((bp + 1) * 32 * wc * cpt + rt * (bp + 1)) * es;
      const int64_t tiles = cdiv(rows, rt);
      // CTAs per (excitation, chunk): with plenty of row tiles ~8 waves of CTAs (the hardware scheduler balances the
      // SMs) that each walk >= 32 tiles, so staging D_i stays a few per cent of a CTA's work; small problems fill
      // one wave instead
      const int64_t per = (int64_t)nbatch * nz;
      const int64_t slots = 148 * 2;      // 126 registers x 256 threads: two CTAs per SM
      const int64_t g_fill = std::max<int64_t>(1, slots / per), g_waves = std::max<int64_t>(1, (slots * (int64_t)g_ru_waves) / per);
      int64_t groups = std::max<int64_t>(1, std::min<int64_t>(tiles, std::max(g_fill, std::min(g_waves, tiles / 32))));
      {
        // wave quantisation: 336 CTAs on 296 slots ran as two waves (MESH_50_50: 61 us; 288 CTAs: one wave), 912 as
        // four (200x200).  Among the group counts down to 2/3 of the target take the one whose last wave is fullest.
        auto ctas = [&](int64_t g) { return per * (int64_t)cdiv(tiles, cdiv(tiles, g)); };
        auto eff = [&](int64_t g) { const int64_t c = ctas(g); return (double)c / ((double)cdiv(c, slots) * slots); };
        int64_t best_g = groups;
        double best_e = eff(groups);
        for (int64_t g = groups - 1; g >= std::max<int64_t>(1, (2 * groups) / 3); --g)
          if (eff(g) > best_e + 0.02) { best_e = eff(g); best_g = g; }
        groups = best_g;
      }
      const int passes = (int)std::max<int64_t>(1, cdiv(tiles, groups));
      dim3 grid(nbatch, cdiv(tiles, passes), nz);
#define CEIT_RU(c, w) pdl(rank_update_kernel<T, c, w>, grid, 256, smu, su)(rows, (int)nU, bp, src, yv, Dcat, p->d_ue_ptr, \
                                                                        i0, zero_rows, out, stride, passes, cw)
#define CEIT_RU_W(w) \
  if (cpt == 1) CEIT_RU(1, w); else if (cpt == 2) CEIT_RU(2, w); else CEIT_RU(3, w)
      if (wc == 1) { CEIT_RU_W(1); } else if (wc == 2) { CEIT_RU_W(2); } else { CEIT_RU_W(4); }
#undef CEIT_RU_W
#undef CEIT_RU
      CEIT_COUNT_LAUNCH();
    };
    update(st, nU, Tinv, te, 1, Sinv, u2);
    // the big materialisation Yh_i (st) runs beside ZW and the baseline potentials (sx), which only need Sinv
    if (g_multi_stream && g_ex_overlap) {
      if (int rc_ = ensure_aux_stream(p)) return rc_;
      sx = p->s2;
      CEIT_AFTER(st, sx);
    }
    update(st, N, YT, ye, 0, Yh, N * nU);
    // ZW_i = Xh Dcat_i^T (N x (bp+1)): the rank-(b+1) correction of the element blocks (woodbury_lr_kernel)
    // all excitations of the chunk in one product: Dcat is a (nbatch (bp+1)) x nU matrix
    // (leading dimension rounded up to even: 16-byte aligned rows for the bulk copies of the persistent Woodbury kernel)
    gemm<T>(sx, 0, 1, N, (int64_t)nbatch * (bp + 1), nU, one, Xh, nU, 0, Dcat, nU, 0, zero, (T*)p->ZW,
            zw_ld(nbatch, bp), 0, 1);
  } else {
  dim3 g2(cdiv(u2, 256), nbatch);
  pdl(build_stilde_kernel<T>, g2, 256, 0, st)((int)nU, SU, p->d_inB, i0, St);
  CEIT_COUNT_LAUNCH();
  CUDA_TRY(cudaMemsetAsync(Ls, 0, nbatch * u2 * es, st));
  CUDA_TRY(cudaMemsetAsync(Linvs, 0, nbatch * u2 * es, st));
  potrf_trtri_blocked<T>(st, nU, St, nU, u2, Ls, nU, u2, Linvs, nU, u2, tmpS, (int64_t)TILE * nU, nbatch, p->d_info);
  gemm<T>(st, 1, 0, nU, nU, nU, one, Linvs, nU, u2, Linvs, nU, u2, zero, Sinv, nU, u2, nbatch);
  pdl(mask_sinv_kernel<T>, g2, 256, 0, st)((int)nU, p->d_inB, i0, Sinv);
  CEIT_COUNT_LAUNCH();
  // Yh[b] = Xh Sinv[b]
  gemm<T>(st, 0, 0, N, nU, nU, one, Xh, nU, 0, Sinv, nU, u2, zero, Yh, nU, N * nU, nbatch);
  }
  dim3 g3(cdiv(nU, 8), nbatch);
  // (one CTA per excitation running the four stages back to back was tried: 80 us slower at nU = 264 - a single SM
  //  streams the nU x nU matrices at its load latency; the four chip-wide launches stay)
  pdl(phi_u_kernel<T>, g3, 256, 0, sx)((int)nU, SU, Sinv, p->d_inB, i0, tvec, phi0, N, N0, 0);
  CEIT_COUNT_LAUNCH();
  pdl(phi_u_kernel<T>, g3, 256, 0, sx)((int)nU, SU, Sinv, p->d_inB, i0, tvec, phi0, N, N0, 1);
  CEIT_COUNT_LAUNCH();
  pdl(phi_u_kernel<T>, g3, 256, 0, sx)((int)nU, SU, Sinv, p->d_inB, i0, tvec, phi0, N, N0, 2);
  CEIT_COUNT_LAUNCH();
  pdl(phi_u_kernel<T>, g3, 256, 0, sx)((int)nU, SU, Sinv, p->d_inB, i0, tvec, phi0, N, N0, 3);
  CEIT_COUNT_LAUNCH();
  if (nbatch >= 4) {
    // phi_F0 = -X phi_U for all excitations of the chunk as ONE product (X is read once, not once per excitation):
    // Phi^T (nbatch x N0) = -Phi_U^T (nbatch x nU, the tail of each phi0 row) X^T
    gemm<T>(sx, 0, 1, nbatch, N0, nU, neg(one), phi0 + N0, N, 0, Xh, nU, 0, zero, phi0, N, 0, 1);
  } else {
    dim3 g4(cdiv(N0, 8), nbatch);
    pdl(phi_f_kernel<T>, g4, 256, 0, sx)(N0, (int)nU, Xh, phi0, N);
    CEIT_COUNT_LAUNCH();
  }
  CEIT_AFTER(sx, st);
  KERNEL_CHECK();
  return 0;
}

int excitation_chunk_cap(const ceit_plan* p, size_t es) {
  const HostPlan& h = p->h;
  const double budget = 24.0e9;  // bytes of Yh kept at once
  double per = (double)h.Nc * h.nU * es + 4.0 * h.nU * h.nU * es;
  int cap = (int)std::max(1.0, std::min((double)h.L, budget / per));
  return cap;
}

int launch_woodbury_patch(ceit_plan* p, const double* g0e, const double* Xh, const double* Yh, const double* phi0,
                          int e_from, int e_to, int nbatch, int i0, double s_coef, double* J, int64_t ldJ,
                          cudaStream_t st) {
  const HostPlan& h = p->h;
  const size_t psm = woodbury_patch_smem(h.nU, h.patch_node_cap);
  double4* consts = (double4*)p->tmpS;      // free after the batched factorisation of the excitation stage
  pdl(node_consts_kernel, dim3(cdiv(h.nU, 128), nbatch), 128, 0, st)((int)h.nU, h.Nc, h.N0c, phi0, p->d_wk, consts);
  CEIT_COUNT_LAUNCH();
  CUDA_TRY(cudaFuncSetAttribute(woodbury_patch_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psm));
  dim3 grid((unsigned)((int64_t)p->n_patches * nbatch));
  pdl(woodbury_patch_kernel, grid, WBP_WARPS * 32, psm, st)(e_from, e_to, (int)h.L, (int)h.nU, (int)woodbury_pitch(h.nU),
                                                          h.Nc, h.N0c, p->d_patch_eptr, p->d_patch_elems, p->d_patch_nptr,
                                                          p->d_patch_nodes, p->d_patch_local, p->d_area, g0e, Xh, Yh,
                                                          phi0, p->d_ue_ptr, consts, i0, s_coef, J, ldJ,
                                                          (int)h.patch_node_cap, nbatch);
  CEIT_COUNT_LAUNCH();
  return 0;
}
// per-excitation constants of the low-rank kernel (launched outside the Woodbury stage timer, so that the timed
// region of stage 3 is the Woodbury kernel alone - the duration bench.py's roofline uses)
int launch_woodbury_lr_consts(ceit_plan* p, int nbatch, cudaStream_t st) {
  const HostPlan& h = p->h;
  const int nUp = (int)((h.nU + 1) & ~1LL), cst_len = (int)woodbury_lr_cst_len(h.nU, h.L);
  double* consts = (double*)p->tmpS;        // free after the batched factorisation of the excitation stage
  pdl(node_consts_soa_kernel, dim3(cdiv(std::max(h.nU, h.L), 128), nbatch), 128, 0, st)(
      (int)h.nU, nUp, (int)h.L, cst_len, h.Nc, h.N0c, (const double*)p->phi0, p->d_wk, p->d_ue_ptr, consts);
  CEIT_COUNT_LAUNCH();
  return 0;
}
int launch_woodbury_lr(ceit_plan* p, int e_from, int e_to, int nbatch, int i0, double s_coef, double* J, int64_t ldJ,
                       cudaStream_t st) {
  const HostPlan& h = p->h;
  const size_t psm = woodbury_lr_smem(h.nU, h.patch_node_cap, h.bp, h.L);
  const double* consts = (const double*)p->tmpS;
  CUDA_TRY(cudaFuncSetAttribute(woodbury_lr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psm));
  dim3 grid((unsigned)((int64_t)p->n_patches * nbatch));
  pdl(woodbury_lr_kernel, grid, WBP_WARPS * 32, psm, st)(
      e_from, e_to, (int)h.L, (int)h.nU, woodbury_lr_pitch((int)h.nU), (int)h.bp, h.Nc, h.N0c, p->d_patch_eptr,
      p->d_patch_elems, p->d_patch_nptr, p->d_patch_nodes, p->d_patch_local, p->d_area, (const double*)p->g1e,
      (const double*)p->YT, (const double*)p->ye, (const double*)p->ZW, (const double*)p->Yh, (const double*)p->phi0,
      p->d_ue_ptr, consts, i0, s_coef, J, ldJ, (int)h.patch_node_cap, nbatch, zw_ld(nbatch, h.bp));
  CEIT_COUNT_LAUNCH();
  return 0;
}

int launch_woodbury_lr_persist(ceit_plan* p, int e_from, int e_to, int nbatch, int i0, double s_coef, double* J,
                               int64_t ldJ, cudaStream_t st) {
  const HostPlan& h = p->h;
  const size_t psm = (size_t)wbq_smem_bytes(p->wbq.wmax, (int)h.L, h.patch_node_cap, (int)h.bp);
  const double* consts = (const double*)p->tmpS;
  if (p->sm_count == 0) {
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaDeviceGetAttribute(&p->sm_count, cudaDevAttrMultiProcessorCount, dev));
  }
  CUDA_TRY(cudaFuncSetAttribute(woodbury_lr_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psm));
  const int64_t n_items = (int64_t)p->n_patches * nbatch * p->wbq.n;
  if (n_items == 0) return 0;
  dim3 grid((unsigned)std::min<int64_t>(p->sm_count, n_items));
  pdl(woodbury_lr_persist_kernel, grid, WBQ_THREADS, psm, st)(
      e_from, e_to, (int)h.L, (int)h.nU, (int)h.bp, h.Nc, h.N0c, (int)n_items, p->wbq, p->d_patch_eptr,
      p->d_patch_nptr, p->d_patch_nodes, p->d_patch_rec, (const double*)p->g1e, (const double*)p->YT,
      (const double*)p->ye, (const double*)p->ZW, (const double*)p->Yh, (const double*)p->phi0, p->d_ue_ptr, consts, i0,
      s_coef, J, ldJ, (int)h.patch_node_cap, nbatch, zw_ld(nbatch, h.bp));
  CEIT_COUNT_LAUNCH();
  return 0;
}

int launch_woodbury_patch(ceit_plan*, const cd*, const cd*, const cd*, const cd*, int, int, int, int, double, double*,
                          int64_t, cudaStream_t) {
  return CEIT_ERR_STATE;   // the patch kernel is the real-base path; complex uses woodbury_kernel<cd>
}

template <class T>
int jacobian_impl(ceit_plan* p, int64_t i_from, int64_t i_to, int64_t e_from, int64_t e_to, double c, double omega,
                  double* J, int64_t ldJ, double* base, cudaStream_t st) {
  const HostPlan& h = p->h;
  const int cap = std::min<int64_t>(excitation_chunk_cap(p, sizeof(T)), i_to - i_from);
  int rc = alloc_excitation_ws<T>(p, cap, st);
  if (rc) return rc;
  const int warps = 8;
  const size_t smem = (size_t)warps * h.nU * sizeof(double);
  if (smem > 48 * 1024)
    CUDA_TRY(cudaFuncSetAttribute(woodbury_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  for (int64_t i0 = i_from; i0 < i_to; i0 += cap) {
    const int nbatch = (int)std::min<int64_t>(cap, i_to - i0);
    {
      StageScope sc(2, st);
      rc = excitation_impl<T>(p, (int)i0, nbatch, st);
    }
    if (rc) return rc;
    if (base) {
      pdl(baseline_kernel<T>, nbatch, 64, 0, st)((int)h.L, (int)h.nU, h.Nc, h.N0c, (const T*)p->phi0, p->d_we_ptr,
                                                p->d_we_u, p->d_we_w, (int)i0, base, nullptr);
      CEIT_COUNT_LAUNCH();
    }
    const bool lowrank = use_rank_update(p) && !g_no_lowrank &&
                         woodbury_lr_smem(h.nU, h.patch_node_cap, h.bp, h.L) <= (size_t)WB_SMEM_SM;
    const bool use_patch = !g_no_patch && !Sc<T>::is_complex && h.patch_node_cap >= 12 && h.u_contiguous &&
                           (lowrank || woodbury_patch_smem(h.nU, h.patch_node_cap) <= (size_t)WB_SMEM_SM);
    if (e_to > e_from && J && use_patch) {
      if (lowrank && (rc = launch_woodbury_lr_consts(p, nbatch, st))) return rc;
      StageScope sc(3, st);
      if (lowrank && p->wbq_ok && g_wb_persist)
        rc = launch_woodbury_lr_persist(p, (int)e_from, (int)e_to, nbatch, (int)i0, 2.0 * omega * c, J, ldJ, st);
      else if (lowrank)
        rc = launch_woodbury_lr(p, (int)e_from, (int)e_to, nbatch, (int)i0, 2.0 * omega * c, J, ldJ, st);
      else
        rc = launch_woodbury_patch(p, (const T*)p->g0e, (const T*)p->Xh, (const T*)p->Yh, (const T*)p->phi0, (int)e_from,
                                   (int)e_to, nbatch, (int)i0, 2.0 * omega * c, J, ldJ, st);
      if (rc) return rc;
    } else if (e_to > e_from && J) {
      StageScope sc(3, st);
      dim3 grid(cdiv(e_to - e_from, warps), nbatch);
      pdl(woodbury_kernel<T>, grid, warps * 32, smem, st)((int)e_from, (int)e_to, (int)h.L, (int)h.nU, h.Nc, h.N0c,
                                                         p->d_elem_pos, p->d_area, (const T*)p->g0e, (const T*)p->Xh,
                                                         (const T*)p->Yh, (const T*)p->phi0, p->d_we_ptr, p->d_we_u,
                                                         p->d_we_w, (int)i0, 2.0 * omega * c, J, ldJ);
      CEIT_COUNT_LAUNCH();
    }
    KERNEL_CHECK();
  }
  return 0;
}

template <class T>
int forward_impl(ceit_plan* p, int64_t i, double* node_abs, double* elem_pot, double* electrode_pot,
                 cudaStream_t st) {
  const HostPlan& h = p->h;
  int rc = alloc_excitation_ws<T>(p, std::max(1, p->ex_cap), st);
  if (rc) return rc;
  rc = excitation_impl<T>(p, (int)i, 1, st);
  if (rc) return rc;
  if (electrode_pot) {
    pdl(baseline_kernel<T>, 1, 64, 0, st)((int)h.L, (int)h.nU, h.Nc, h.N0c, (const T*)p->phi0, p->d_we_ptr, p->d_we_u,
                                         p->d_we_w, (int)i, nullptr, electrode_pot);
    CEIT_COUNT_LAUNCH();
  }
  if (node_abs || elem_pot) {
    double* na = node_abs;
    if (!na) {
      if (!p->d_node_abs) {
        rc = dev_alloc(p, (void**)&p->d_node_abs, h.N * sizeof(double));
        if (rc) return rc;
      }
      na = p->d_node_abs;
    }
    pdl(node_abs_kernel<T>, cdiv(h.N, 256), 256, 0, st)(h.N, (const T*)p->phi0, p->d_pos, na);
    CEIT_COUNT_LAUNCH();
    if (elem_pot) {
      pdl(elem_mean_kernel, cdiv(h.E, 256), 256, 0, st)(h.E, p->d_elem, na, elem_pot);
      CEIT_COUNT_LAUNCH();
    }
  }
  KERNEL_CHECK();
  return 0;
}

}  // namespace

namespace {
// Stream-ordered temporaries of the plan-less inverse-model entry points.  They come from a pool OWNED by the
// library (its release threshold keeps the memory across synchronisations; the device's default pool, which
// PyTorch and other users share, is left alone) and are released on every exit path.
cudaMemPool_t inverse_pool() {
  static cudaMemPool_t pool = nullptr;
  static int pool_dev = -1;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  if (pool && pool_dev == dev) return pool;
  cudaMemPoolProps props;
  std::memset(&props, 0, sizeof(props));
  props.allocType = cudaMemAllocationTypePinned;
  props.handleTypes = cudaMemHandleTypeNone;
  props.location.type = cudaMemLocationTypeDevice;
  props.location.id = dev;
  cudaMemPool_t np = nullptr;
  if (cudaMemPoolCreate(&np, &props) != cudaSuccess) return nullptr;
  uint64_t thr = UINT64_MAX;
  cudaMemPoolSetAttribute(np, cudaMemPoolAttrReleaseThreshold, &thr);
  pool = np;
  pool_dev = dev;
  return pool;
}
struct AsyncBufs {
  cudaStream_t st;
  std::vector<void*> bufs;
  explicit AsyncBufs(cudaStream_t s) : st(s) {}
  ~AsyncBufs() {
    for (void* b : bufs) cudaFreeAsync(b, st);
  }
  cudaError_t get(void** out, size_t bytes) {
    if (bytes == 0) bytes = 16;
    cudaMemPool_t pool = inverse_pool();
    cudaError_t e = pool ? cudaMallocFromPoolAsync(out, bytes, pool, st) : cudaMallocAsync(out, bytes, st);
    if (e == cudaSuccess) bufs.push_back(*out);
    return e;
  }
  template <class V>
  cudaError_t get(V** out, int64_t count) { return get((void**)out, (size_t)count * sizeof(V)); }
};
}  // namespace

// ====================================================================================================
extern "C" {

const char* ceit_last_error(void) { return ceit::g_err.c_str(); }
int ceit_version(void) { return 100; }
int64_t ceit_launch_count(void) { return ceit::g_launches.load(); }

int ceit_plan_create(const double* nodes, const int64_t* elem, int64_t N, int64_t E, const int64_t* el_ptr,
                     const int64_t* el_elems, int64_t L, int64_t target_block, int upload, ceit_plan_t** out) {
  if (!nodes || !elem || !el_ptr || !el_elems || !out) CEIT_FAIL(CEIT_ERR_ARG, "ceit_plan_create: null argument");
  ceit_plan* p = new ceit_plan();
  try {
    build_host_plan(p->h, nodes, elem, N, E, el_ptr, el_elems, L, target_block, g_nd_interior,
                    g_nd_tree != 0 && target_block <= 0, g_nd_pull, g_nd_ways, g_shard_parts, g_shard_part);
    {
      // patches of <= 32 elements whose distinct node rows fit the shared-memory budget
      const bool lr = rank_update_shape_ok(p->h) && !g_direct_excitation && !g_no_lowrank;
      int64_t cap = g_patch_nodes > 0 ? std::min(g_patch_nodes, 64)
                                      : (lr ? woodbury_lr_cap(p->h.nU, p->h.bp, p->h.L) : woodbury_patch_cap(p->h.nU));
      int qcap = 0;
      if (lr && g_wb_persist && g_patch_nodes <= 0 && wbq_configure(p->h, &p->wbq, &qcap)) {
        p->wbq_ok = true;
        cap = qcap;
      }
      if (cap < 3) cap = 3;
      build_patches(p->h, (int32_t)cap, 32);
    }
  } catch (const std::exception& e) {
    delete p;
    CEIT_FAIL(CEIT_ERR_ARG, e.what());
  }
  if (upload) {
    int rc = upload_plan(p);
    if (rc) {
      ceit_plan_destroy(p);
      return rc;
    }
  }
  *out = p;
  return CEIT_OK;
}

int ceit_plan_destroy(ceit_plan_t* p) {
  if (!p) return CEIT_OK;
  if (p->s2) {
    cudaStreamSynchronize(p->s2);
    cudaStreamDestroy(p->s2);
    for (cudaStream_t s : {p->s3, p->s4, p->s5})
      if (s) {
        cudaStreamSynchronize(s);
        cudaStreamDestroy(s);
      }
    p->lane_c.destroy();
    p->lane_a.destroy();
    p->lane_b.destroy();
    for (auto e : p->ev_ring)
      if (e) cudaEventDestroy(e);
  }
  for (void* a : p->allocs)
    if (a) cudaFree(a);
  delete p;
  return CEIT_OK;
}

int ceit_plan_info(const ceit_plan_t* p, int64_t* info) {
  if (!p || !info) CEIT_FAIL(CEIT_ERR_ARG, "ceit_plan_info: null argument");
  const HostPlan& h = p->h;
  std::memset(info, 0, 24 * sizeof(int64_t));
  info[0] = h.N; info[1] = h.E; info[2] = h.L; info[3] = h.nU; info[4] = h.N0; info[5] = h.nb;
  info[6] = h.max_block; info[7] = h.nnz; info[8] = h.n_levels; info[9] = h.szD; info[10] = h.szS;
  info[11] = (int64_t)p->dev_bytes; info[12] = p->mode;
  info[13] = 0; info[14] = (int64_t)h.patch_eptr.size() - 1; info[15] = (int64_t)h.patch_nodes.size();
  info[16] = h.P; info[17] = h.nI; info[18] = h.bmax; info[19] = h.nC; info[20] = h.NP; info[21] = h.N0P;
  info[22] = h.patch_node_cap;
  info[23] = h.use_tree ? h.tree.H : 0;
  if (p->uploaded && p->d_info) {
    int flag = 0;
    if (cudaMemcpy(&flag, p->d_info, sizeof(int), cudaMemcpyDeviceToHost) == cudaSuccess) info[13] = flag;
  }
  return CEIT_OK;
}

int64_t ceit_plan_generation(const ceit_plan_t* p) { return p ? p->generation : -1; }

int ceit_plan_numeric_flag(const ceit_plan_t* p, int* flag_out, void* stream) {
  if (!p || !flag_out) CEIT_FAIL(CEIT_ERR_ARG, "ceit_plan_numeric_flag: null argument");
  *flag_out = 0;
  if (!p->uploaded || !p->d_info) return CEIT_OK;
  cudaStream_t st = (cudaStream_t)stream;
  CUDA_TRY(cudaMemcpyAsync(flag_out, p->d_info, sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return CEIT_OK;
}

int ceit_plan_tables(const ceit_plan_t* p, int64_t* rowptr, int64_t* colidx, int64_t* pos, int64_t* blk_ptr,
                     int64_t* U, int64_t* ppos) {
  if (!p) CEIT_FAIL(CEIT_ERR_ARG, "ceit_plan_tables: null plan");
  const HostPlan& h = p->h;
  if (rowptr) for (int64_t k = 0; k <= h.N; ++k) rowptr[k] = h.rowptr[k];
  if (colidx) for (int64_t k = 0; k < h.nnz; ++k) colidx[k] = h.colidx[k];
  if (pos) for (int64_t k = 0; k < h.N; ++k) pos[k] = h.pos[k];
  if (ppos) for (int64_t k = 0; k < h.N; ++k) ppos[k] = h.ppos[k];
  if (blk_ptr) for (int64_t k = 0; k <= h.nb; ++k) blk_ptr[k] = h.blk_ptr[k];
  if (U) for (int64_t k = 0; k < h.nU; ++k) U[k] = h.U[k];
  return CEIT_OK;
}

int ceit_plan_patches(const ceit_plan_t* p, int64_t* eptr, int64_t* elems, int64_t* nptr, int64_t* nodes,
                      int64_t* local) {
  if (!p) CEIT_FAIL(CEIT_ERR_ARG, "ceit_plan_patches: null plan");
  const HostPlan& h = p->h;
  if (eptr) for (size_t k = 0; k < h.patch_eptr.size(); ++k) eptr[k] = h.patch_eptr[k];
  if (elems) for (size_t k = 0; k < h.patch_elems.size(); ++k) elems[k] = h.patch_elems[k];
  if (nptr) for (size_t k = 0; k < h.patch_nptr.size(); ++k) nptr[k] = h.patch_nptr[k];
  if (nodes) for (size_t k = 0; k < h.patch_nodes.size(); ++k) nodes[k] = h.patch_nodes[k];
  if (local) for (size_t k = 0; k < h.patch_local.size(); ++k) local[k] = h.patch_local[k];
  return CEIT_OK;
}

int ceit_assemble(ceit_plan_t* p, const double* perm, const double* var, double omega, double* vals,
                  double* elem_param, void* stream) {
  if (!p || !perm || !var) CEIT_FAIL(CEIT_ERR_ARG, "ceit_assemble: null argument");
  if (!p->uploaded) CEIT_FAIL(CEIT_ERR_STATE, "ceit_assemble: plan has no device tables (created with upload=0)");
  cudaStream_t st = (cudaStream_t)stream;
  const HostPlan& h = p->h;
  StageScope sc(0, st);
  if (g_assemble_gather) {
    pdl(assemble_csr_kernel, cdiv(h.nnz, 256), 256, 0, st)((int)h.nnz, p->d_src_ptr, p->d_src, p->d_nodes, p->d_elem,
                                                          perm, var, omega, p->d_vals);
    CEIT_COUNT_LAUNCH();
  } else {
    pdl(assemble_elem_kernel, cdiv(h.E, 256), 256, 0, st)((int)h.E, p->d_nodes, p->d_elem, perm, var, omega,
                                                         p->d_elem_dst, (double2*)p->d_contrib);
    CEIT_COUNT_LAUNCH();
    pdl(assemble_reduce_kernel, cdiv(h.nnz, 256), 256, 0, st)((int)h.nnz, p->d_src_ptr, (const double2*)p->d_contrib,
                                                             p->d_vals);
    CEIT_COUNT_LAUNCH();
  }
  KERNEL_CHECK();
  if (vals) CUDA_TRY(cudaMemcpyAsync(vals, p->d_vals, h.nnz * sizeof(cd), cudaMemcpyDeviceToDevice, st));
  if (elem_param) {
    pdl(elem_param_kernel, cdiv(h.E, 256), 256, 0, st)((int)h.E, p->d_nodes, p->d_elem, elem_param, nullptr);
    CEIT_COUNT_LAUNCH();
    KERNEL_CHECK();
  }
  p->assembled = true;
  p->mode = 0;
  return CEIT_OK;
}

int ceit_plan_csr_values(const ceit_plan_t* p, double* vals, void* stream) {
  if (!p || !vals) CEIT_FAIL(CEIT_ERR_ARG, "ceit_plan_csr_values: null argument");
  if (!p->assembled) CEIT_FAIL(CEIT_ERR_STATE, "ceit_plan_csr_values: call ceit_assemble first");
  CUDA_TRY(cudaMemcpyAsync(vals, p->d_vals, p->h.nnz * sizeof(cd), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return CEIT_OK;
}

int ceit_factor(ceit_plan_t* p, int is_complex, void* stream) {
  if (!p) CEIT_FAIL(CEIT_ERR_ARG, "ceit_factor: null plan");
  if (!p->assembled) CEIT_FAIL(CEIT_ERR_STATE, "ceit_factor: call ceit_assemble first");
  cudaStream_t st = (cudaStream_t)stream;
  if (p->h.tree.parts > 1 && p->h.use_tree)
    CEIT_FAIL(CEIT_ERR_STATE, "ceit_factor: a domain-decomposed plan is factorised with ceit_factor_phase(1), the exchange, ceit_factor_phase(2)");
  StageScope sc(1, st);
  return is_complex ? factor_impl<cd>(p, st) : factor_impl<double>(p, st);
}

int ceit_factor_phase(ceit_plan_t* p, int phase, void* stream) {
  if (!p) CEIT_FAIL(CEIT_ERR_ARG, "ceit_factor_phase: null plan");
  if (!p->assembled) CEIT_FAIL(CEIT_ERR_STATE, "ceit_factor_phase: call ceit_assemble first");
  if (!p->h.use_tree || p->h.tree.parts <= 1) CEIT_FAIL(CEIT_ERR_STATE, "ceit_factor_phase: the plan is not domain-decomposed");
  if (phase != 1 && phase != 2) CEIT_FAIL(CEIT_ERR_ARG, "ceit_factor_phase: phase must be 1 or 2");
  if (!use_rank_update(p)) CEIT_FAIL(CEIT_ERR_STATE, "ceit_factor_phase: the decomposed path needs the rank-update excitation stage");
  cudaStream_t st = (cudaStream_t)stream;
  StageScope sc(1, st);
  return factor_impl<double>(p, st, phase);
}

int ceit_plan_exchange(const ceit_plan_t* p, int64_t* info, void** gather_buf, void** reduce_buf) {
  if (!p || !info) CEIT_FAIL(CEIT_ERR_ARG, "ceit_plan_exchange: null argument");
  const HostPlan& h = p->h;
  const TreePlan& t = h.tree;
  info[0] = t.parts; info[1] = t.part;
  info[2] = (int64_t)t.roots_per_rank * t.slot_sz;                       // doubles per rank in the all-gather buffer
  info[3] = h.use_tree ? (t.rowsA - t.ra_top0 + h.nU) * h.nU : 0;         // doubles in the all-reduce region
  info[4] = h.Nc; info[5] = h.N0c; info[6] = t.dcut; info[7] = t.n_own_segs;
  if (gather_buf) *gather_buf = p->xg;
  if (reduce_buf) *reduce_buf = (p->K0U && h.use_tree) ? (void*)((double*)p->K0U + t.ra_top0 * h.nU) : nullptr;
  return CEIT_OK;
}

int ceit_plan_own_elements(const ceit_plan_t* p, uint8_t* own) {
  if (!p || !own) CEIT_FAIL(CEIT_ERR_ARG, "ceit_plan_own_elements: null argument");
  for (int64_t e = 0; e < p->h.E; ++e) own[e] = p->h.elem_own[e];
  return CEIT_OK;
}

int ceit_jacobian_block(ceit_plan_t* p, int64_t i_from, int64_t i_to, int64_t e_from, int64_t e_to, double c,
                        double omega, double* J, int64_t ldJ, double* base, void* stream) {
  if (!p) CEIT_FAIL(CEIT_ERR_ARG, "ceit_jacobian_block: null plan");
  if (p->mode <= 0) CEIT_FAIL(CEIT_ERR_STATE, "ceit_jacobian_block: call ceit_factor first");
  const HostPlan& h = p->h;
  if (i_from < 0 || i_to > h.L || i_from > i_to || e_from < 0 || e_to > h.E || e_from > e_to)
    CEIT_FAIL(CEIT_ERR_ARG, "ceit_jacobian_block: range out of bounds");
  if (J && ldJ < h.E) CEIT_FAIL(CEIT_ERR_ARG, "ceit_jacobian_block: ldJ < E");
  if (i_from == i_to) return CEIT_OK;
  cudaStream_t st = (cudaStream_t)stream;
  return p->mode == 2 ? jacobian_impl<cd>(p, i_from, i_to, e_from, e_to, c, omega, J, ldJ, base, st)
                      : jacobian_impl<double>(p, i_from, i_to, e_from, e_to, c, omega, J, ldJ, base, st);
}

int ceit_forward(ceit_plan_t* p, int64_t i, double* node_abs, double* elem_pot, double* electrode_pot,
                 void* stream) {
  if (!p) CEIT_FAIL(CEIT_ERR_ARG, "ceit_forward: null plan");
  if (p->mode <= 0) CEIT_FAIL(CEIT_ERR_STATE, "ceit_forward: call ceit_factor first");
  if (i < 0 || i >= p->h.L) CEIT_FAIL(CEIT_ERR_ARG, "ceit_forward: electrode index out of range");
  if (p->h.tree.parts > 1 && p->h.use_tree)
    CEIT_FAIL(CEIT_ERR_STATE, "ceit_forward: node potentials of the whole mesh need a plan without domain decomposition");
  cudaStream_t st = (cudaStream_t)stream;
  return p->mode == 2 ? forward_impl<cd>(p, i, node_abs, elem_pot, electrode_pot, st)
                      : forward_impl<double>(p, i, node_abs, elem_pot, electrode_pot, st);
}

int ceit_inverse_model(const double* J, int64_t ldJ, int64_t m, const int64_t* det_idx, int64_t n_det,
                       const int64_t* rows, int64_t m_sel, double shift, double lambda, int form, double scale_,
                       double* inv_out, int* flag, void* stream) {
  if (!J || !det_idx || !inv_out) CEIT_FAIL(CEIT_ERR_ARG, "ceit_inverse_model: null argument");
  if (m <= 0 || n_det <= 0 || ldJ <= 0) CEIT_FAIL(CEIT_ERR_ARG, "ceit_inverse_model: bad shape");
  if (!rows) m_sel = m;
  if (m_sel <= 0) CEIT_FAIL(CEIT_ERR_ARG, "ceit_inverse_model: empty row selection");
  cudaStream_t st = (cudaStream_t)stream;
  StageScope sc(4, st);
  if (form == 0) form = (m_sel <= n_det) ? 2 : 1;
  const int64_t n = (form == 2) ? m_sel : n_det;  // size of the SPD system
  double *Jt = nullptr, *G = nullptr, *Lg = nullptr, *Li = nullptr, *Gi = nullptr, *tmp = nullptr;
  int* info = flag;
  AsyncBufs ws(st);
  CUDA_TRY(ws.get(&Jt, m_sel * n_det));
  CUDA_TRY(ws.get(&G, n * n));
  CUDA_TRY(ws.get(&Lg, n * n));
  CUDA_TRY(ws.get(&Li, n * n));
  CUDA_TRY(ws.get(&Gi, std::max(n * n, n_det * m_sel)));
  CUDA_TRY(ws.get(&tmp, (int64_t)TILE * n));
  if (!info) CUDA_TRY(ws.get(&info, 1));
  CUDA_TRY(cudaMemsetAsync(info, 0, sizeof(int), st));
  CUDA_TRY(cudaMemsetAsync(Lg, 0, n * n * sizeof(double), st));
  CUDA_TRY(cudaMemsetAsync(Li, 0, n * n * sizeof(double), st));
  pdl(gather_shift_kernel, cdiv(m_sel * n_det, 256), 256, 0, st)(m_sel, n_det, J, ldJ, rows, det_idx, shift, Jt);
  CEIT_COUNT_LAUNCH();
  if (form == 2) {
    // G = Jt Jt^T + lambda^2 I (m x m);  inv = Jt^T G^-1
    {
      const int S = splitk_slices(n, n, n_det, 1);
      double* sws = nullptr;
      if (S > 1) CUDA_TRY(ws.get(&sws, (int64_t)S * n * n));
      if (S > 1) gemm_splitk(st, 0, 1, n, n, n_det, 1.0, Jt, n_det, Jt, n_det, 0.0, G, n, sws, /*lower_only=*/1);
      else gemm<double>(st, 0, 1, n, n, n_det, 1.0, Jt, n_det, 0, Jt, n_det, 0, 0.0, G, n, 0, 1, /*lower_only=*/1);
    }
    pdl(add_diag_kernel, cdiv(n, 256), 256, 0, st)(n, lambda * lambda, G, n);
    CEIT_COUNT_LAUNCH();
    potrf_trtri_blocked<double>(st, n, G, n, 0, Lg, n, 0, Li, n, 0, tmp, 0, 1, info, inv_lane());
    gemm<double>(st, 1, 0, n, n, n, 1.0, Li, n, 0, Li, n, 0, 0.0, Gi, n, 0);
    gemm<double>(st, 1, 0, n_det, m_sel, m_sel, scale_, Jt, n_det, 0, Gi, n, 0, 0.0, inv_out, m_sel, 0);
  } else {
    // A = Jt^T Jt + lambda^2 I (n_det x n_det); inv = Linv^T (Linv Jt^T)
    gemm<double>(st, 1, 0, n, n, m_sel, 1.0, Jt, n_det, 0, Jt, n_det, 0, 0.0, G, n, 0, 1, /*lower_only=*/1);
    pdl(add_diag_kernel, cdiv(n, 256), 256, 0, st)(n, lambda * lambda, G, n);
    CEIT_COUNT_LAUNCH();
    potrf_trtri_blocked<double>(st, n, G, n, 0, Lg, n, 0, Li, n, 0, tmp, 0, 1, info, inv_lane());
    gemm<double>(st, 0, 1, n, m_sel, n, 1.0, Li, n, 0, Jt, n_det, 0, 0.0, Gi, m_sel, 0);
    gemm<double>(st, 1, 0, n, m_sel, n, scale_, Li, n, 0, Gi, m_sel, 0, 0.0, inv_out, m_sel, 0);
  }
  KERNEL_CHECK();
  return CEIT_OK;
}

// ---- sharded inverse model (dual form): partial Gram matrix, then rows of inv from the reduced Gram ----
int ceit_gram_partial(const double* J, int64_t ldJ, int64_t m, const int64_t* det_idx, int64_t n_det,
                      const int64_t* rows, int64_t m_sel, double shift, double* G, void* stream) {
  if (!J || !det_idx || !G) CEIT_FAIL(CEIT_ERR_ARG, "ceit_gram_partial: null argument");
  if (!rows) m_sel = m;
  if (m <= 0 || n_det < 0 || m_sel <= 0) CEIT_FAIL(CEIT_ERR_ARG, "ceit_gram_partial: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  StageScope sc(4, st);
  if (n_det == 0) {
    CUDA_TRY(cudaMemsetAsync(G, 0, m_sel * m_sel * sizeof(double), st));
    return CEIT_OK;
  }
  double* Jt = nullptr;
  AsyncBufs ws(st);
  CUDA_TRY(ws.get(&Jt, m_sel * n_det));
  pdl(gather_shift_kernel, cdiv(m_sel * n_det, 256), 256, 0, st)(m_sel, n_det, J, ldJ, rows, det_idx, shift, Jt);
  CEIT_COUNT_LAUNCH();
  {
    // symmetric: lower triangle only (split over K when its tiles cannot fill the chip), then mirrored so that the
    // all-reduce over the ranks sums a complete matrix
    const int S = splitk_slices(m_sel, m_sel, n_det, 1);
    double* sws = nullptr;
    if (S > 1) CUDA_TRY(ws.get(&sws, (int64_t)S * m_sel * m_sel));
    gemm_splitk(st, 0, 1, m_sel, m_sel, n_det, 1.0, Jt, n_det, Jt, n_det, 0.0, G, m_sel, sws, /*lower_only=*/1);
    pdl(symmetrize_from_lower_kernel, cdiv(m_sel * m_sel, 256), 256, 0, st)(m_sel, G, m_sel);
    CEIT_COUNT_LAUNCH();
  }
  KERNEL_CHECK();
  return CEIT_OK;
}

int ceit_inverse_from_gram(const double* J, int64_t ldJ, int64_t m, const int64_t* det_idx, int64_t n_det,
                           const int64_t* rows, int64_t m_sel, double shift, double lambda, double scale_,
                           const double* G, double* inv_out, int* flag, void* stream) {
  if (!J || !det_idx || !G || !inv_out) CEIT_FAIL(CEIT_ERR_ARG, "ceit_inverse_from_gram: null argument");
  if (!rows) m_sel = m;
  if (m <= 0 || n_det <= 0 || m_sel <= 0) CEIT_FAIL(CEIT_ERR_ARG, "ceit_inverse_from_gram: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  StageScope sc(4, st);
  const int64_t n = m_sel;
  double *Jt = nullptr, *Gw = nullptr, *Lg = nullptr, *Li = nullptr, *Gi = nullptr, *tmp = nullptr;
  int* info = flag;
  AsyncBufs ws(st);
  CUDA_TRY(ws.get(&Jt, m_sel * n_det));
  CUDA_TRY(ws.get(&Gw, n * n));
  CUDA_TRY(ws.get(&Lg, n * n));
  CUDA_TRY(ws.get(&Li, n * n));
  CUDA_TRY(ws.get(&Gi, n * n));
  CUDA_TRY(ws.get(&tmp, (int64_t)TILE * n));
  if (!info) CUDA_TRY(ws.get(&info, 1));
  CUDA_TRY(cudaMemsetAsync(info, 0, sizeof(int), st));
  CUDA_TRY(cudaMemsetAsync(Lg, 0, n * n * sizeof(double), st));
  CUDA_TRY(cudaMemsetAsync(Li, 0, n * n * sizeof(double), st));
  CUDA_TRY(cudaMemcpyAsync(Gw, G, n * n * sizeof(double), cudaMemcpyDeviceToDevice, st));
  pdl(add_diag_kernel, cdiv(n, 256), 256, 0, st)(n, lambda * lambda, Gw, n);
  CEIT_COUNT_LAUNCH();
  pdl(gather_shift_kernel, cdiv(m_sel * n_det, 256), 256, 0, st)(m_sel, n_det, J, ldJ, rows, det_idx, shift, Jt);
  CEIT_COUNT_LAUNCH();
  potrf_trtri_blocked<double>(st, n, Gw, n, 0, Lg, n, 0, Li, n, 0, tmp, 0, 1, info, inv_lane());
  gemm<double>(st, 1, 0, n, n, n, 1.0, Li, n, 0, Li, n, 0, 0.0, Gi, n, 0);
  gemm<double>(st, 1, 0, n_det, m_sel, m_sel, scale_, Jt, n_det, 0, Gi, n, 0, 0.0, inv_out, m_sel, 0);
  KERNEL_CHECK();
  return CEIT_OK;
}

int ceit_reconstruct(const double* inv, const double* dV, double* out, int64_t n_det, int64_t m, int64_t F,
                     void* stream) {
  if (!inv || !dV || !out) CEIT_FAIL(CEIT_ERR_ARG, "ceit_reconstruct: null argument");
  if (n_det <= 0 || m <= 0 || F <= 0) CEIT_FAIL(CEIT_ERR_ARG, "ceit_reconstruct: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  StageScope sc(5, st);
  if (F == 1) {
    pdl(gemv_rows_kernel, cdiv(n_det, 8), 256, 0, st)(n_det, (int)m, inv, dV, out);
    CEIT_COUNT_LAUNCH();
  } else {
    gemm<double>(st, 0, 0, n_det, F, m, 1.0, inv, m, 0, dV, F, 0, 0.0, out, F, 0);
  }
  KERNEL_CHECK();
  return CEIT_OK;
}

int ceit_mesh_model(const double* nodes, const int64_t* elem, int64_t N, int64_t E, const double* centers, int64_t L,
                    double radius, double det_bound, double* elem_param, uint8_t* in_electrode, uint8_t* in_detection,
                    void* stream) {
  if (!nodes || !elem || (in_electrode && !centers)) CEIT_FAIL(CEIT_ERR_ARG, "ceit_mesh_model: null argument");
  if (N <= 0 || E <= 0 || L < 0) CEIT_FAIL(CEIT_ERR_ARG, "ceit_mesh_model: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  pdl(mesh_model_kernel, cdiv(E, 256), 256, 0, st)(E, (int)L, nodes, elem, centers, radius, det_bound, elem_param,
                                                  in_electrode, in_detection);
  CEIT_COUNT_LAUNCH();
  KERNEL_CHECK();
  return CEIT_OK;
}

int ceit_center_of_position(const double* maps, int64_t n_det, int64_t F, const double* cx, const double* cy,
                            double frac, double* out, void* stream) {
  if (!maps || !cx || !cy || !out) CEIT_FAIL(CEIT_ERR_ARG, "ceit_center_of_position: null argument");
  if (n_det <= 0 || F <= 0) CEIT_FAIL(CEIT_ERR_ARG, "ceit_center_of_position: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  pdl(cop_frames_kernel, cdiv(F, 32), 256, 0, st)(n_det, F, maps, cx, cy, frac, out);
  CEIT_COUNT_LAUNCH();
  KERNEL_CHECK();
  return CEIT_OK;
}

int ceit_dgemm(int opA, int opB, int64_t M, int64_t N, int64_t K, double alpha, const double* A, int64_t lda,
               const double* B, int64_t ldb, double beta, double* C, int64_t ldc, void* stream) {
  if (!A || !B || !C) CEIT_FAIL(CEIT_ERR_ARG, "ceit_dgemm: null argument");
  gemm<double>((cudaStream_t)stream, opA, opB, M, N, K, alpha, A, lda, 0, B, ldb, 0, beta, C, ldc, 0);
  KERNEL_CHECK();
  return CEIT_OK;
}

int ceit_zgemm(int opA, int opB, int64_t M, int64_t N, int64_t K, double alpha_re, double alpha_im, const double* A,
               int64_t lda, const double* B, int64_t ldb, double beta_re, double beta_im, double* C, int64_t ldc,
               void* stream) {
  if (!A || !B || !C) CEIT_FAIL(CEIT_ERR_ARG, "ceit_zgemm: null argument");
  gemm<cd>((cudaStream_t)stream, opA, opB, M, N, K, mk(alpha_re, alpha_im), (const cd*)A, lda, 0, (const cd*)B, ldb, 0,
           mk(beta_re, beta_im), (cd*)C, ldc, 0);
  KERNEL_CHECK();
  return CEIT_OK;
}

int ceit_stage_times(double* ms, int64_t* counts) {
  if (!ms || !counts) CEIT_FAIL(CEIT_ERR_ARG, "ceit_stage_times: null argument");
  for (int k = 0; k < N_STAGES; ++k) {
    ms[k] = 0.0;
    counts[k] = 0;
  }
  for (auto& e : g_stage_events) {
    CUDA_TRY(cudaEventSynchronize(e.b));
    float t = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&t, e.a, e.b));
    ms[e.stage] += t;
    counts[e.stage] += 1;
    cudaEventDestroy(e.a);
    cudaEventDestroy(e.b);
  }
  g_stage_events.clear();
  return CEIT_OK;
}

int ceit_stage_flops(double* flops) {
  if (!flops) CEIT_FAIL(CEIT_ERR_ARG, "ceit_stage_flops: null argument");
  for (int k = 0; k < N_STAGES; ++k) {
    flops[k] = g_stage_flops[k];
    g_stage_flops[k] = 0.0;
  }
  return CEIT_OK;
}

int ceit_debug_wb_clocks(int64_t* out) {   // out[16]: CTA 81 (first wave) in [0,8), CTA 1500 in [8,16)
  long long h[16];
  CUDA_TRY(cudaMemcpyFromSymbol(h, ceit::g_wb_clk, sizeof(h)));
  for (int k = 0; k < 16; ++k) out[k] = h[k];
  return CEIT_OK;
}

int ceit_debug_wbq_clocks(int64_t* out) {  // out[512]: see WBQ_CLK in kernels.cuh
  static long long h[512];
  CUDA_TRY(cudaMemcpyFromSymbol(h, ceit::g_wbq_clk, sizeof(h)));
  for (int k = 0; k < 512; ++k) out[k] = h[k];
  return CEIT_OK;
}

int ceit_debug_tile_clocks(int64_t* out) {
  long long h[8];
  CUDA_TRY(cudaMemcpyFromSymbol(h, ceit::g_tile_clk, sizeof(h)));
  for (int k = 0; k < 8; ++k) out[k] = h[k];
  return CEIT_OK;
}

int ceit_set_option(const char* name, int64_t value) {
  if (!name) CEIT_FAIL(CEIT_ERR_ARG, "ceit_set_option: null name");
  if (std::strcmp(name, "force_simt_gemm") == 0) {
    ceit::g_force_simt = (int)value;
    return CEIT_OK;
  }
  if (std::strcmp(name, "direct_excitation") == 0) {
    g_direct_excitation = (int)value;
    return CEIT_OK;
  }
  if (std::strcmp(name, "nd_pull") == 0) { g_nd_pull = (int)value; return CEIT_OK; }
  if (std::strcmp(name, "shard_parts") == 0) { g_shard_parts = (int)std::max<int64_t>(1, value); return CEIT_OK; }
  if (std::strcmp(name, "shard_part") == 0) { g_shard_part = (int)value; return CEIT_OK; }
  if (std::strcmp(name, "pdl") == 0) { g_pdl = (int)value; return CEIT_OK; }
  if (std::strcmp(name, "wb_persist") == 0) { g_wb_persist = (int)value; return CEIT_OK; }
  if (std::strcmp(name, "assemble_gather") == 0) { g_assemble_gather = (int)value; return CEIT_OK; }
  if (std::strcmp(name, "yt_sweep") == 0) { g_yt_sweep = (int)value; return CEIT_OK; }
  if (std::strcmp(name, "front_kernel") == 0) { g_front_kernel = (int)value; return CEIT_OK; }
  if (std::strcmp(name, "nd_ways") == 0) { g_nd_ways = (int)value; return CEIT_OK; }
  if (std::strcmp(name, "nd_tree") == 0) {
    g_nd_tree = (int)value;
    return CEIT_OK;
  }
  if (std::strcmp(name, "nd_interior") == 0) {
    g_nd_interior = value;
    return CEIT_OK;
  }
  if (std::strcmp(name, "multi_stream") == 0) {
    g_multi_stream = (int)value;
    return CEIT_OK;
  }
  if (std::strcmp(name, "excitation_overlap") == 0) { g_ex_overlap = (int)value; return CEIT_OK; }
  if (std::strcmp(name, "gemm_min_ctas64") == 0) { ceit::g_gemm_min_ctas64 = (int)value; return CEIT_OK; }
  if (std::strcmp(name, "ru_wc") == 0) { g_ru_wc = (int)value; return CEIT_OK; }
  if (std::strcmp(name, "ru_cpt") == 0) { g_ru_cpt = (int)value; return CEIT_OK; }
  if (std::strcmp(name, "ru_waves") == 0) { g_ru_waves = (int)std::max<int64_t>(1, value); return CEIT_OK; }
  if (std::strcmp(name, "patch_nodes") == 0) {
    g_patch_nodes = (int)value;
    return CEIT_OK;
  }
  if (std::strcmp(name, "no_lowrank_kernel") == 0) {
    g_no_lowrank = (int)value;
    return CEIT_OK;
  }
  if (std::strcmp(name, "no_patch_kernel") == 0) {
    g_no_patch = (int)value;
    return CEIT_OK;
  }
  if (std::strcmp(name, "gemm_vec") == 0) {
    ceit::g_gemm_vec = (int)value;
    return CEIT_OK;
  }
  if (std::strcmp(name, "gemm_tile") == 0) {
    ceit::g_gemm_tile = (int)value;
    return CEIT_OK;
  }
  if (std::strcmp(name, "stage_timers") == 0) {
    g_stage_timers = (int)value;
    return CEIT_OK;
  }
  CEIT_FAIL(CEIT_ERR_ARG, std::string("ceit_set_option: unknown option ") + name);
}

}  // extern "C"
