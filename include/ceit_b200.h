/* ceit_b200 - C ABI of the B200-native CEIT Jacobian / Gauss-Newton hot path.
 *
 * The reference (zehao99/CEIT) is pure Python/numpy and has no FFI; the seam it offers is its
 * Python class surface.  Every entry point below names the reference method whose arithmetic it
 * replaces (file:line relative to the reference root).  The Python drop-in classes in
 * `ceit_b200/` (EFEM, EJAC, Solver) call these through ctypes; INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error; `ceit_last_error()` gives the message
 *     (thread-local, valid until the next failing call on that thread);
 *   - pointers marked [host] are host memory, [dev] are device memory owned by the caller
 *     (torch tensors used as plain buffers); nothing is retained past the call except what the
 *     opaque plan owns (freed by ceit_plan_destroy);
 *   - `stream` is a cudaStream_t passed as void*; work is enqueued on it, no hidden sync unless
 *     the function is documented as `_host` (those synchronise the stream before returning);
 *   - matrices are row-major (C / numpy order); complex values are interleaved (re, im) f64;
 *   - there is NO CPU fallback: every compute entry fails with CEIT_ERR_CUDA if no device.
 */
#ifndef CEIT_B200_H
#define CEIT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CEIT_OK 0
#define CEIT_ERR_ARG (-1)
#define CEIT_ERR_CUDA (-2)
#define CEIT_ERR_STATE (-3)
#define CEIT_ERR_NUMERIC (-4)

typedef struct ceit_plan ceit_plan_t;

const char* ceit_last_error(void);
int ceit_version(void);
/* number of this library's kernels launched by the calling process so far (bench `gpu_launches`) */
int64_t ceit_launch_count(void);

/* ---- plan: symbolic analysis of mesh + electrodes (host), static tables uploaded -----------
 * Replaces the bookkeeping the reference redoes on every forward solve:
 *   EFEM.set_boundary_condition / swap (CEIT/EFEM.py:71-96,136-142), FEMBasic.calc_init
 *   (CEIT/FEMBasic.py:105-114) and the node->element->electrode averaging lists
 *   (CEIT/FEMBasic.py:118-137).
 * nodes [host] f64[N*2]; elem [host] i64[E*3] (0-based, MeshObj.elem, CEIT/models/mesh.py:31);
 * el_ptr [host] i64[L+1], el_elems [host] i64[el_ptr[L]]: electrode_mesh (mesh.py:90-113).
 * target_block: minimum nodes per diagonal block of the level-set block-tridiagonal ordering of the
 * separator system (<=0: default); ceit_set_option("nd_interior", n) sets the target size of the
 * nested-dissection interiors (<= 0: no interiors, every node is a level-2 node).  Works without a GPU when `upload` is 0 (host tables only, for CPU tests). */
int ceit_plan_create(const double* nodes, const int64_t* elem, int64_t N, int64_t E,
                     const int64_t* el_ptr, const int64_t* el_elems, int64_t L,
                     int64_t target_block, int upload, ceit_plan_t** out);
int ceit_plan_destroy(ceit_plan_t* plan);

/* info[0..23]: N, E, L, nU, N0, n_blocks, max_block, nnz, n_levels, sum n_k^2, sum n_k n_k+1,
 *              device bytes held, factor state (0 none,1 real,2 complex), numeric flag, n_patches,
 *              patch nodes, P (interiors), nI (slots per interior), bmax, nC (separator nodes),
 *              NP (position rows), N0P (position of the first electrode node), node cap of a Woodbury patch,
 *              levels of the multi-level nested dissection (0: interiors + block-tridiagonal chain) */
int ceit_plan_info(const ceit_plan_t* plan, int64_t* info24);
/* Allocation generation of the plan: incremented whenever the numeric or per-excitation workspaces are
 * (re)allocated (first use, real <-> complex switch, larger excitation chunk).  A CUDA graph captured from calls on
 * this plan holds the workspace pointers of the generation it was captured in and must be dropped when the
 * generation changes.  Workspaces are never (re)allocated while the stream is being captured: the call fails with
 * CEIT_ERR_STATE instead (run the sequence eagerly once first).  No device access. */
int64_t ceit_plan_generation(const ceit_plan_t* plan);
/* Pivot flag of the factorisations enqueued on `stream` so far (0 ok, 1 = a Cholesky pivot was not positive /
 * zero): copied on `stream` and synchronised there, so it is ordered after the work the caller enqueued on that
 * stream (info[13] of ceit_plan_info reads the same flag with a blocking legacy-stream copy).
 * Mirrors the LinAlgError np.linalg.inv raises on a singular K_ff (CEIT/EFEM.py:119). */
int ceit_plan_numeric_flag(const ceit_plan_t* plan, int* flag_out, void* stream);
/* host copies of the symbolic tables (any pointer may be NULL):
 * rowptr i64[N+1], colidx i64[nnz] (CSR pattern of K, sorted, == scipy.sparse.csr_matrix of the
 * reference's dense K_sparse), pos i64[N] (node -> row of Xh / Yh / phi0: a permutation, F0 nodes
 * first, electrode node u at N0 + u), blk_ptr i64[n_blocks+1] (level-2 local), U i64[nU],
 * ppos i64[N] (node -> padded row of the factor stage: interior p slot a -> p*nI + a, separator nodes
 * in level-2 block order from P*nI, electrode nodes from N0P). */
int ceit_plan_tables(const ceit_plan_t* plan, int64_t* rowptr, int64_t* colidx, int64_t* pos,
                     int64_t* blk_ptr, int64_t* U, int64_t* ppos);

/* host copies of the element patches of the Woodbury kernel (any pointer may be NULL); sizes from
 * ceit_plan_info: n_patches = info[14], total patch nodes = info[15], node cap per patch = info[22]:
 * eptr i64[n_patches+1] / elems i64[E] (patch -> element ids, lane order), nptr i64[n_patches+1] /
 * nodes i64[info[15]] (patch -> node POSITIONS, in shared-memory slot order), local i64[3E] (per patch
 * element, in patch order: the slot of each of its three nodes).  No reference counterpart: the patches replace
 * the element loop of EJAC.JAC_calculation (CEIT/EJAC.py:115-133). */
int ceit_plan_patches(const ceit_plan_t* plan, int64_t* eptr, int64_t* elems, int64_t* nptr, int64_t* nodes,
                      int64_t* local);

/* ---- K1: element matrices -> CSR values ------------------------------------------------------
 * Replaces EFEM.construct_sparse_matrix (CEIT/EFEM.py:47-69) incl. its uniform factor 2 and the
 * per-element geometry of MeshObj.initialize_parameters (CEIT/models/mesh.py:41-75).
 * perm [dev] f64[E] (= mesh.elem_perm, already scaled by 1/resistance, FEMBasic.py:57-58),
 * var [dev] f64[E] (= elem_variable), omega = 2*pi*signal_frequency (FEMBasic.py:60).
 * vals [dev] complex f64[nnz] (may be NULL: only the internal block storage is filled).
 * elem_param [dev] f64[E*9] optional output: [A,b0,b1,b2,c0,c1,c2,xbar,ybar] (mesh.py:70-71). */
int ceit_assemble(ceit_plan_t* plan, const double* perm, const double* var, double omega,
                  double* vals, double* elem_param, void* stream);

/* copy the CSR values of the last ceit_assemble into vals [dev] complex f64[nnz] */
int ceit_plan_csr_values(const ceit_plan_t* plan, double* vals, void* stream);

/* ---- K3: base stage (factor once per material state) ------------------------------------------
 * Replaces the np.linalg.inv(K_ff) of EVERY forward solve (CEIT/EFEM.py:111-119): block Cholesky
 * of K00 = K[F0,F0], X = K00^-1 K0U, S_U, selected inverse on the mesh pattern.
 * is_complex: 0 = real path (requires var == 0 everywhere), 1 = complex symmetric path. */
int ceit_factor(ceit_plan_t* plan, int is_complex, void* stream);

/* ---- K3 on several GPUs: domain-decomposed base stage ------------------------------------------------
 * The reference's only parallel axis is the excitation range calc_from / calc_end (CEIT/EJAC.py:107-119); here the
 * MESH is decomposed instead (north_star: "the element loop shards by element range"): a plan created after
 * ceit_set_option("shard_parts", G) / ("shard_part", r) cuts the elimination forest of K00 at the smallest depth with
 * >= G sub-trees; rank r factorises the sub-trees of its region, every rank the shared top, and everything downstream
 * (X, Y_T, the per-excitation updates, the Woodbury kernel) runs on the rank's own node rows and elements only - so
 * the base stage, the excitation stage and the element loop all shrink with G, and the inverse model is fed by the
 * rank's own Jacobian columns (ceit_gram_partial / ceit_inverse_from_gram; J itself is only gathered for output).
 * Sequence per material state (real path):
 *     ceit_assemble; ceit_factor_phase(plan, 1, s);
 *     [caller, NCCL on stream s]  all-gather(gather_buf: info[2] doubles per rank, in place),
 *                                 all-reduce-sum(reduce_buf: info[3] doubles)
 *     ceit_factor_phase(plan, 2, s); ceit_jacobian_block(...)   -> columns of ceit_plan_own_elements only
 * ceit_plan_exchange: info i64[8] = {G, r, doubles per rank of gather_buf, doubles of reduce_buf, rows held by this
 * rank (Nc), of which non-electrode (N0c), cut depth, rank-owned segments}; the buffer pointers [dev] are valid
 * after the first ceit_factor_phase(1) and until the plan is destroyed or changes between real and complex. */
int ceit_factor_phase(ceit_plan_t* plan, int phase, void* stream);
int ceit_plan_exchange(const ceit_plan_t* plan, int64_t* info, void** gather_buf, void** reduce_buf);
/* own [host] u8[E]: 1 = this rank computes the element's Jacobian columns (all 1 without decomposition) */
int ceit_plan_own_elements(const ceit_plan_t* plan, uint8_t* own);

/* ---- K2: Jacobian block ------------------------------------------------------------------------
 * Replaces EJAC.JAC_calculation's double loop (CEIT/EJAC.py:115-133): for excitations
 * [i_from,i_to) (calc_from/calc_end semantics) and elements [e_from,e_to): perturb element j by
 * c = variable_change_for_JAC, solve, store |electrode_potential[m]|, m != i ascending, at
 * J[(i*(L-1)+cnt) * ldJ + j].  J [dev] f64, row-major, ldJ >= E.
 * base [dev] f64[L*(L-1)] optional: electrode_original_potential (EJAC.calc_origin_potential,
 * CEIT/EJAC.py:87-98) rows of the same excitation range.
 * Requires ceit_assemble + ceit_factor on the current material state. */
int ceit_jacobian_block(ceit_plan_t* plan, int64_t i_from, int64_t i_to, int64_t e_from,
                        int64_t e_to, double c, double omega, double* J, int64_t ldJ,
                        double* base, void* stream);

/* ---- single forward solve -----------------------------------------------------------------------
 * Replaces EFEM.calculation(i) (CEIT/FEMBasic.py:67-86; EFEM.py:37-45,98-134):
 * node_abs [dev] f64[N] = |phi| per node in ORIGINAL node order, elem_pot [dev] f64[E],
 * electrode_pot [dev] f64[L] (any may be NULL). Requires assemble + factor. */
int ceit_forward(ceit_plan_t* plan, int64_t i, double* node_abs, double* elem_pot,
                 double* electrode_pot, void* stream);

/* ---- inverse model -----------------------------------------------------------------------------
 * Replaces Solver.get_inv_matrix / EJAC.save_inv_matrix (CEIT/Solver.py:110-124,
 * CEIT/EJAC.py:231-243) including eliminate_non_detect_JAC (Solver.py:76-86) and the "-1":
 *   Jt = J[:, det] - shift ;  inv = (Jt^T Jt + lambda^2 I)^-1 Jt^T      (n_det x m)
 * J [dev] f64[m*ldJ]; det_idx [dev] i64[n_det]; rows [dev] i64[m_sel] optional row subset
 * (eit_solve_on_some_electrodes, EJAC.py:210-220; NULL = all m rows);
 * form: 0 = auto, 1 = primal (n_det x n_det normal equations, as the reference writes it),
 *       2 = dual (m x m push-through form, identical operator); scale multiplies the result
 * (1e-4 in EJAC.py:228).  inv_out [dev] f64[n_det * m_sel].
 * flag [dev] int, optional: set to 0, then to 1 if a pivot of the Cholesky factorisation of the regularised
 * normal matrix is not positive (NaN / Inf in J, lambda = 0 with a rank-deficient Gram matrix) - the condition on
 * which the reference's np.linalg.inv raises LinAlgError (Solver.py:121); the host classes check it before they
 * cache inv_mat.npy. */
int ceit_inverse_model(const double* J, int64_t ldJ, int64_t m, const int64_t* det_idx,
                       int64_t n_det, const int64_t* rows, int64_t m_sel, double shift,
                       double lambda, int form, double scale, double* inv_out, int* flag, void* stream);

/* ---- inverse model sharded over detection elements (multi-GPU, dual form) -------------------------
 * Each rank holds a slice det_idx[0..n_det) of the detection elements (J replicated):
 *   ceit_gram_partial      G_r = Jt_r Jt_r^T (m_sel x m_sel) over its slice; the caller sums G over ranks
 *                          (one all-reduce of m_sel^2 doubles);
 *   ceit_inverse_from_gram rows of inv for the slice: Jt_r^T (G + lambda^2 I)^-1  (n_det x m_sel).
 * Same arithmetic as ceit_inverse_model(form = 2) (CEIT/Solver.py:110-124). */
int ceit_gram_partial(const double* J, int64_t ldJ, int64_t m, const int64_t* det_idx, int64_t n_det,
                      const int64_t* rows, int64_t m_sel, double shift, double* G, void* stream);
int ceit_inverse_from_gram(const double* J, int64_t ldJ, int64_t m, const int64_t* det_idx, int64_t n_det,
                           const int64_t* rows, int64_t m_sel, double shift, double lambda, double scale,
                           const double* G, double* inv_out, int* flag, void* stream);

/* ---- realtime reconstruction ---------------------------------------------------------------------
 * Replaces Solver.solve (CEIT/Solver.py:95-108): out = inv_mat . delta_V
 * inv [dev] f64[n_det*m]; dV [dev] f64[m*F] row-major (F=1: a vector); out [dev] f64[n_det*F]. */
int ceit_reconstruct(const double* inv, const double* dV, double* out, int64_t n_det, int64_t m,
                     int64_t F, void* stream);

/* ---- mesh model on the device ------------------------------------------------------------------------------
 * Replaces the per-element Python loops of MeshObj.initialize_parameters (3x3 np.linalg.inv per element,
 * CEIT/models/mesh.py:41-75), calc_electrode_elements (mesh.py:90-113) and calc_detection_elements (mesh.py:127-155).
 * nodes [dev] f64[N*2]; elem [dev] i64[E*3]; centers [dev] f64[L*2] and radius: the electrode squares (metres);
 * det_bound: half-width of the detection square.  Outputs (each optional): elem_param [dev] f64[E*9] =
 * [A, b0,b1,b2, c0,c1,c2, xbar, ybar]; in_electrode [dev] u8[L*E]: 1 when the centroid of element e lies in the CLOSED
 * square of electrode m (row m); in_detection [dev] u8[E]: 1 when |xbar| < det_bound and |ybar| < det_bound.  The
 * ascending index lists the reference keeps (electrode_mesh, detection_index) are the non-zero positions. */
int ceit_mesh_model(const double* nodes, const int64_t* elem, int64_t N, int64_t E, const double* centers, int64_t L,
                    double radius, double det_bound, double* elem_param, uint8_t* in_electrode, uint8_t* in_detection,
                    void* stream);

/* ---- consumer of the reconstructed maps: centre of position -------------------------------------------
 * Replaces DataPlotter.get_COP (CEIT/datahandler/DataHandler.py:28-56), batched over frames:
 * threshold = frac * max + (1 - frac) * min of a frame (the reference's THRESHOLD_PERCENTAGE = 0.9), then the
 * mean of (cx, cy, value) over the detection elements strictly above it.
 * maps [dev] f64[n_det*F] row-major (what ceit_reconstruct writes); cx, cy [dev] f64[n_det]: centroid coordinates
 * of the detection elements (elem_param[detection_index][7|8], CEIT/models/mesh.py:70-71);
 * out [dev] f64[F*4] = x_mean, y_mean, v_mean, count per frame (count = 0 gives NaN means, where the reference
 * raises ZeroDivisionError). */
int ceit_center_of_position(const double* maps, int64_t n_det, int64_t F, const double* cx, const double* cy,
                            double frac, double* out, void* stream);

/* ---- generic dense helper exposed for tests/benchmarks: C = alpha op(A) op(B) + beta C ---------
 * row-major f64, op: 0 = N, 1 = T. Uses the same DMMA kernel as the factorisation. */
int ceit_dgemm(int opA, int opB, int64_t M, int64_t N, int64_t K, double alpha, const double* A,
               int64_t lda, const double* B, int64_t ldb, double beta, double* C, int64_t ldc,
               void* stream);

/* complex (interleaved re, im) variant: lda/ldb/ldc in complex elements; non-conjugating transposes */
int ceit_zgemm(int opA, int opB, int64_t M, int64_t N, int64_t K, double alpha_re, double alpha_im,
               const double* A, int64_t lda, const double* B, int64_t ldb, double beta_re, double beta_im,
               double* C, int64_t ldc, void* stream);

/* per-stage device times (CUDA events on the caller's stream) accumulated since the last call, when
 * ceit_set_option("stage_timers", 1) is on.  ms f64[8], counts i64[8]; stages: 0 assemble, 1 base
 * factorisation, 2 per-excitation stage, 3 Woodbury kernel, 4 inverse model, 5 reconstruct.
 * Synchronises on the recorded events. */
int ceit_stage_times(double* ms, int64_t* counts);
/* algorithmic real flops of the dense work (DMMA GEMMs: 2 M N K, half for lower-only symmetric results, 8 M N K
 * complex; tile factorisation + triangular inverse: 2/3 n^3) enqueued inside each stage's timed regions since the last
 * call (same stage indices, same switch as ceit_stage_times): the numerator of the per-stage FP64 roofline.
 * flops f64[8].  Host-side count, no device access. */
int ceit_stage_flops(double* flops);

/* debug: clock64 phase stamps of the last tile-factorisation CTA (only with -DCEIT_TILE_CLOCKS) */
int ceit_debug_tile_clocks(int64_t* out8);
/* debug: phase stamps of the Woodbury kernels (only with -DCEIT_TILE_CLOCKS): out16 = two sample CTAs of the
 * one-CTA-per-item kernel; out512 = per item of CTA 7 of the persistent kernel (tools/wb_clocks.py, wbq_clocks.py) */
int ceit_debug_wb_clocks(int64_t* out16);
int ceit_debug_wbq_clocks(int64_t* out512);

/* switches (value 0/1 unless noted):
 *   "stage_timers"      record CUDA events around every stage (ceit_stage_times); not capturable
 *   "multi_stream"      (default 1) independent chains of a stage run on the plan's helper streams and
 *                       join before the call returns its work to `stream`; 0 = everything on `stream`
 *   "nd_interior"       target nodes per nested-dissection interior for plans created afterwards (default 64)
 *   "patch_nodes"       node cap of the element patches for plans created afterwards (0 = automatic)
 *   "gemm_tile"         force the CTA tile of the DMMA GEMM (32 / 64 / 128, 0 = automatic)
 *   "gemm_vec"          (default 1) stage the GEMM operands with 16-byte cp.async when base / leading dimension /
 *                       batch stride are even; 0 = always 8-byte copies
 *   "nd_ways"           (default 4) plans created afterwards: cuts per front of the nested dissection - 4 = both halves
 *                       of a sub-domain are cut again and the three separators form one front with up to four children
 *                       (half as many levels on the serial spine), 2 = plain bisection
 *   "nd_pull"           (default 1) plans created afterwards: 1 = every level pulls its children's update matrices in
 *                       one destination-centric launch, 0 = children push (one extend-add pass per sibling index)
 *   "front_kernel"      (default 0) 1: pull levels of small fronts (own <= 64, boundary <= 64) run pull +
 *                       factorisation + A^-1 / Y / U / W as ONE kernel per level (measured slower on MESH_50_50: one
 *                       SM per front); the complex path always runs the separate steps
 *   "yt_sweep"          -1 (default): Y_T = Xh T by a second backward sweep over the elimination forest when the
 *                       N x nU x nU product would exceed 2e10 flops; 0 / 1 force the product / the sweep (set before
 *                       the first ceit_factor of a plan: the sweep needs two extra N0P x nU buffers)
 *   "nd_tree"           (default 1) multi-level nested dissection of the base stage for plans created afterwards;
 *                       0 = one level of interiors + block-tridiagonal separator chain (round-1 ordering)
 *   "gemm_min_ctas64"   automatic tile choice: 64x64 tiles when they give at least this many CTAs, else 32x32
 *                       (default 300)
 *   "ru_wc", "ru_cpt", "ru_waves"
 *                       rank-update kernel: force the column layout (warps across x columns per thread; 0 =
 *                       cost model) and the number of CTA waves (default 8); tools/ru_sweep.sh
 *   "shard_parts", "shard_part"
 *                       plans created afterwards: domain decomposition over G ranks, this process being rank r
 *                       (default 1 / 0 = none); see ceit_factor_phase
 *   "pdl"               (default 1) programmatic dependent launch: 1 = the DMMA GEMM and tile-factorisation kernels are
 *                       launched with the programmatic-stream-serialisation attribute (they wait for their predecessor
 *                       with griddepcontrol.wait after their prologue; in a captured graph the edges become
 *                       programmatic), 2 = every kernel of the path, 0 = plain launches
 *   "wb_persist"        (default 1) low-rank Woodbury stage as ONE persistent CTA per SM walking (patch, excitation,
 *                       column chunk) items through a 3-slot TMA ring with two consumer groups (even nU); 0 = one CTA
 *                       per (patch, excitation)
 *   "assemble_gather"   (default 0) 1: K1 as one thread per CSR slot re-deriving the triangle geometry of each of its
 *                       contributions (round-1 form); 0 = element-parallel pass (geometry once per element, stores
 *                       into the slot-sorted contribution list) + segmented sum per CSR slot
 *   "excitation_overlap" (default 1) ZW / base potentials on a second stream beside the Yh update
 *   "force_simt_gemm", "direct_excitation", "no_patch_kernel", "no_lowrank_kernel"
 *                       route around the tensor-core GEMM / rank-update excitation stage / patch Woodbury
 *                       kernels (test coverage of the fallback kernels)
 *
 * CUDA graphs: after ONE eager call sequence (the library allocates its workspaces on first use), any
 * sequence of ceit_assemble / ceit_factor / ceit_jacobian_block / ceit_forward / ceit_inverse_model /
 * ceit_reconstruct calls on one stream may be captured with cudaStreamBeginCapture and replayed: no entry
 * point synchronises, the helper streams fork from and join back into `stream` with events, and the
 * workspace of ceit_inverse_model comes from cudaMallocAsync (allocation nodes).  Replaying the ~140 short
 * dependent kernels of a MESH_50_50 step from a graph removes ~0.2 ms of launch gaps (bench.py does). */
int ceit_set_option(const char* name, int64_t value);

#ifdef __cplusplus
}
#endif
#endif
