"""ctypes binding of libceit_b200.so (the C ABI in include/ceit_b200.h).

PyTorch is used only for device buffers and the current stream; every numeric step on the path is
a kernel of this library.  There is NO CPU fallback: if the shared library or a CUDA device is
missing, the calls below raise.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libceit_b200.so")
_lib = None

SYMBOLS = [
    "ceit_last_error", "ceit_version", "ceit_launch_count", "ceit_plan_create", "ceit_plan_destroy",
    "ceit_plan_info", "ceit_plan_tables", "ceit_plan_patches", "ceit_assemble", "ceit_factor", "ceit_jacobian_block",
    "ceit_forward", "ceit_inverse_model", "ceit_reconstruct", "ceit_dgemm", "ceit_plan_csr_values",
    "ceit_set_option", "ceit_stage_times", "ceit_gram_partial", "ceit_inverse_from_gram", "ceit_debug_tile_clocks", "ceit_zgemm",
    "ceit_center_of_position", "ceit_plan_generation", "ceit_plan_numeric_flag", "ceit_stage_flops", "ceit_mesh_model",
    "ceit_factor_phase", "ceit_plan_exchange", "ceit_plan_own_elements",
]


class CeitError(RuntimeError):
    pass


def lib():
    """Load (building first if the .so is absent and nvcc is available) the native library."""
    global _lib
    if _lib is not None:
        return _lib
    from . import build as _build
    if not _build.is_current():          # absent, or older than the sources (content hash): never load a stale library
        try:
            _build.build()
        except Exception as e:  # pragma: no cover
            if not os.path.exists(LIB_PATH):
                raise CeitError(
                    f"native library {LIB_PATH} is missing and could not be built ({e}); "
                    "run `python -m ceit_b200.build` - there is no CPU fallback") from e
            raise CeitError(f"native library {LIB_PATH} is older than its sources and could not be rebuilt ({e}); "
                            "run `python -m ceit_b200.build`") from e
    L = ctypes.CDLL(LIB_PATH)
    c_i64, c_dbl, c_vp, c_int = ctypes.c_int64, ctypes.c_double, ctypes.c_void_p, ctypes.c_int
    L.ceit_last_error.restype = ctypes.c_char_p
    L.ceit_version.restype = c_int
    L.ceit_launch_count.restype = c_i64
    L.ceit_plan_create.argtypes = [c_vp, c_vp, c_i64, c_i64, c_vp, c_vp, c_i64, c_i64, c_int,
                                   ctypes.POINTER(c_vp)]
    L.ceit_plan_destroy.argtypes = [c_vp]
    L.ceit_plan_info.argtypes = [c_vp, c_vp]
    L.ceit_plan_tables.argtypes = [c_vp] * 7
    L.ceit_plan_patches.argtypes = [c_vp] * 6
    L.ceit_assemble.argtypes = [c_vp, c_vp, c_vp, c_dbl, c_vp, c_vp, c_vp]
    L.ceit_factor.argtypes = [c_vp, c_int, c_vp]
    L.ceit_jacobian_block.argtypes = [c_vp, c_i64, c_i64, c_i64, c_i64, c_dbl, c_dbl, c_vp, c_i64, c_vp, c_vp]
    L.ceit_forward.argtypes = [c_vp, c_i64, c_vp, c_vp, c_vp, c_vp]
    L.ceit_inverse_model.argtypes = [c_vp, c_i64, c_i64, c_vp, c_i64, c_vp, c_i64, c_dbl, c_dbl, c_int,
                                     c_dbl, c_vp, c_vp, c_vp]
    L.ceit_reconstruct.argtypes = [c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, c_vp]
    L.ceit_dgemm.argtypes = [c_int, c_int, c_i64, c_i64, c_i64, c_dbl, c_vp, c_i64, c_vp, c_i64, c_dbl,
                             c_vp, c_i64, c_vp]
    L.ceit_set_option.argtypes = [ctypes.c_char_p, c_i64]
    L.ceit_plan_csr_values.argtypes = [c_vp, c_vp, c_vp]
    L.ceit_stage_times.argtypes = [c_vp, c_vp]
    L.ceit_stage_flops.argtypes = [c_vp]
    L.ceit_mesh_model.argtypes = [c_vp, c_vp, c_i64, c_i64, c_vp, c_i64, c_dbl, c_dbl, c_vp, c_vp, c_vp, c_vp]
    L.ceit_gram_partial.argtypes = [c_vp, c_i64, c_i64, c_vp, c_i64, c_vp, c_i64, c_dbl, c_vp, c_vp]
    L.ceit_inverse_from_gram.argtypes = [c_vp, c_i64, c_i64, c_vp, c_i64, c_vp, c_i64, c_dbl, c_dbl, c_dbl, c_vp, c_vp, c_vp, c_vp]
    L.ceit_factor_phase.argtypes = [c_vp, c_int, c_vp]
    L.ceit_plan_exchange.argtypes = [c_vp, c_vp, c_vp, c_vp]
    L.ceit_plan_own_elements.argtypes = [c_vp, c_vp]
    L.ceit_plan_generation.argtypes = [c_vp]
    L.ceit_plan_generation.restype = c_i64
    L.ceit_plan_numeric_flag.argtypes = [c_vp, c_vp, c_vp]
    L.ceit_debug_tile_clocks.argtypes = [c_vp]
    L.ceit_center_of_position.argtypes = [c_vp, c_i64, c_i64, c_vp, c_vp, c_dbl, c_vp, c_vp]
    L.ceit_zgemm.argtypes = [c_int, c_int, c_i64, c_i64, c_i64, c_dbl, c_dbl, c_vp, c_i64, c_vp, c_i64, c_dbl, c_dbl,
                             c_vp, c_i64, c_vp]
    for name in SYMBOLS:
        if name not in ("ceit_last_error", "ceit_launch_count", "ceit_plan_generation"):
            getattr(L, name).restype = c_int
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise CeitError(lib().ceit_last_error().decode("utf-8", "replace") + f" (code {rc})")


STAGES = ["assemble", "factor", "excitation", "woodbury", "inverse_model", "reconstruct", "s6", "s7"]


_options = {}
_replayed_launches = 0


def set_option(name, value):
    check(lib().ceit_set_option(name.encode(), int(value)))
    _options[name] = int(value)


def stage_times():
    """{stage: (total ms, launches)} since the last call (needs set_option('stage_timers', 1))."""
    ms = np.zeros(8, dtype=np.float64)
    cnt = np.zeros(8, dtype=np.int64)
    check(lib().ceit_stage_times(_np_ptr(ms), _np_ptr(cnt)))
    return {STAGES[k]: (float(ms[k]), int(cnt[k])) for k in range(8)}


def stage_flops():
    """{stage: algorithmic real flops of the dense work} since the last call (with stage timers on)."""
    fl = np.zeros(8, dtype=np.float64)
    check(lib().ceit_stage_flops(_np_ptr(fl)))
    return {STAGES[k]: float(fl[k]) for k in range(8)}


def launch_count():
    """Kernels launched by the library so far, counting every replay of a captured StepGraph."""
    return int(lib().ceit_launch_count()) + _replayed_launches


def _np_ptr(a):
    return ctypes.c_void_p(a.ctypes.data)


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise CeitError("ceit_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch


def stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def tptr(t):
    """Device pointer of a torch tensor used as a plain buffer (None -> NULL)."""
    if t is None:
        return ctypes.c_void_p(0)
    assert t.is_cuda and t.is_contiguous()
    return ctypes.c_void_p(t.data_ptr())


class StepGraph:
    """A sequence of library calls on STATIC device buffers, captured once into a CUDA graph and replayed.

    The hot path of one material state is ~130 short dependent kernels; replaying them from a graph removes
    the launch gaps between them (MESH_50_50: 1.56 -> 1.31 ms per step).  `fn` must have run eagerly at least
    once before (the library allocates its workspaces on first use, which is not capturable) and must only
    touch buffers that stay alive and in place: the kernel arguments (pointers, ranges, omega, c) are frozen.
    """

    def __init__(self, fn):
        torch = require_cuda()
        if _options.get("stage_timers", 0):
            raise Exception("StepGraph: stage timers record CUDA events and cannot be captured; switch them off")
        self.graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        n0 = launch_count()
        with torch.cuda.graph(self.graph, stream=side, capture_error_mode="thread_local"):
            self.out = fn()
        self.kernels = launch_count() - n0          # launches recorded per replay
        torch.cuda.current_stream().wait_stream(side)

    def replay(self):
        global _replayed_launches
        self.graph.replay()
        _replayed_launches += self.kernels
        return self.out


class Plan:
    """Opaque `ceit_plan_t*` + the few host tables the Python classes need."""

    def __init__(self, nodes, elem, electrode_lists, target_block=0, upload=True, parts=1, part=0):
        """parts > 1: domain-decomposed plan of rank `part` (include/ceit_b200.h, ceit_factor_phase)."""
        nodes = np.ascontiguousarray(nodes, dtype=np.float64)
        elem = np.ascontiguousarray(elem, dtype=np.int64)
        if nodes.ndim != 2 or nodes.shape[1] != 2:
            raise Exception("2D Node coordinates incorrect")      # FEMBasic.py:49-50
        if elem.ndim != 2 or elem.shape[1] != 3:
            raise Exception("2D Elements dimension incorrect")    # FEMBasic.py:54-55
        ptr = np.zeros(len(electrode_lists) + 1, dtype=np.int64)
        for i, e in enumerate(electrode_lists):
            ptr[i + 1] = ptr[i] + len(e)
        idx = np.ascontiguousarray(np.concatenate([np.asarray(e, dtype=np.int64) for e in electrode_lists])
                                   if len(electrode_lists) else np.zeros(0, np.int64))
        if upload:
            require_cuda()
        self._h = ctypes.c_void_p()
        self.N, self.E, self.L = nodes.shape[0], elem.shape[0], len(electrode_lists)
        self.parts, self.part = int(parts), int(part)
        lib().ceit_set_option(b"shard_parts", self.parts)
        lib().ceit_set_option(b"shard_part", self.part)
        try:
            check(lib().ceit_plan_create(_np_ptr(nodes), _np_ptr(elem), self.N, self.E, _np_ptr(ptr), _np_ptr(idx),
                                         self.L, int(target_block), 1 if upload else 0, ctypes.byref(self._h)))
        finally:
            lib().ceit_set_option(b"shard_parts", 1)
            lib().ceit_set_option(b"shard_part", 0)
        info = self.info()
        self.nU, self.N0, self.n_blocks, self.max_block, self.nnz = (int(info[k]) for k in (3, 4, 5, 6, 7))
        self.P, self.nI, self.bmax, self.nC, self.NP, self.N0P = (int(info[k]) for k in range(16, 22))
        self.tree_levels = int(info[23])

    def info(self):
        out = np.zeros(24, dtype=np.int64)
        check(lib().ceit_plan_info(self._h, _np_ptr(out)))
        return out

    def tables(self):
        rowptr = np.zeros(self.N + 1, np.int64)
        colidx = np.zeros(self.nnz, np.int64)
        pos = np.zeros(self.N, np.int64)
        blk_ptr = np.zeros(self.n_blocks + 1, np.int64)
        U = np.zeros(self.nU, np.int64)
        ppos = np.zeros(self.N, np.int64)
        check(lib().ceit_plan_tables(self._h, _np_ptr(rowptr), _np_ptr(colidx), _np_ptr(pos), _np_ptr(blk_ptr),
                                     _np_ptr(U), _np_ptr(ppos)))
        return dict(rowptr=rowptr, colidx=colidx, pos=pos, blk_ptr=blk_ptr, U=U, ppos=ppos)

    def patches(self):
        """Element patches of the Woodbury kernel (host copies): see ceit_plan_patches in include/ceit_b200.h."""
        info = self.info()
        n_p, n_pn = int(info[14]), int(info[15])
        eptr = np.zeros(n_p + 1, np.int64)
        elems = np.zeros(self.E, np.int64)
        nptr = np.zeros(n_p + 1, np.int64)
        nodes = np.zeros(n_pn, np.int64)
        local = np.zeros(3 * self.E, np.int64)
        check(lib().ceit_plan_patches(self._h, _np_ptr(eptr), _np_ptr(elems), _np_ptr(nptr), _np_ptr(nodes),
                                      _np_ptr(local)))
        return dict(eptr=eptr, elems=elems, nptr=nptr, nodes=nodes, local=local, node_cap=int(info[22]))

    # -- numeric entry points (torch tensors as buffers) ------------------------------------------
    def assemble(self, perm, var, omega, vals=None, elem_param=None):
        check(lib().ceit_assemble(self._h, tptr(perm), tptr(var), float(omega), tptr(vals), tptr(elem_param),
                                  stream_ptr()))

    def csr_values(self, vals):
        check(lib().ceit_plan_csr_values(self._h, tptr(vals), stream_ptr()))

    def factor(self, is_complex):
        check(lib().ceit_factor(self._h, 1 if is_complex else 0, stream_ptr()))

    def factor_phase(self, phase):
        """Half of the base stage of a domain-decomposed plan (1: own sub-trees, 2: after the exchange)."""
        check(lib().ceit_factor_phase(self._h, int(phase), stream_ptr()))

    def exchange_info(self):
        """(info, gather_ptr, reduce_ptr) of ceit_plan_exchange; pointers are device addresses (0 before phase 1)."""
        info = np.zeros(8, np.int64)
        g, r = ctypes.c_void_p(), ctypes.c_void_p()
        check(lib().ceit_plan_exchange(self._h, _np_ptr(info), ctypes.byref(g), ctypes.byref(r)))
        return info, g.value or 0, r.value or 0

    def own_elements(self):
        own = np.zeros(self.E, np.uint8)
        check(lib().ceit_plan_own_elements(self._h, _np_ptr(own)))
        return own.astype(bool)

    def jacobian_block(self, i_from, i_to, e_from, e_to, c, omega, J, ldJ, base=None):
        check(lib().ceit_jacobian_block(self._h, int(i_from), int(i_to), int(e_from), int(e_to), float(c),
                                        float(omega), tptr(J), int(ldJ), tptr(base), stream_ptr()))

    def forward(self, i, node_abs=None, elem_pot=None, electrode_pot=None):
        check(lib().ceit_forward(self._h, int(i), tptr(node_abs), tptr(elem_pot), tptr(electrode_pot),
                                 stream_ptr()))

    def numeric_flag(self):
        """0 = ok, 1 = a Cholesky pivot was not positive (synchronises the current stream)."""
        flag = ctypes.c_int(0)
        check(lib().ceit_plan_numeric_flag(self._h, ctypes.byref(flag), stream_ptr()))
        return int(flag.value)

    def generation(self):
        """Allocation generation of the plan's workspaces (graphs captured in an older one are stale)."""
        return int(lib().ceit_plan_generation(self._h))

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            lib().ceit_plan_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


SINGULAR_MSG = "Singular matrix"      # numpy.linalg.LinAlgError text of np.linalg.inv (Solver.py:121, EJAC.py:165)


def check_flag(flag):
    """Raise like np.linalg.inv does when the regularised normal matrix had a non-positive pivot (synchronises)."""
    if flag is not None and int(flag.item()) != 0:
        raise np.linalg.LinAlgError(SINGULAR_MSG)


def inverse_model(J, det_idx, lmbda, rows=None, shift=1.0, form=0, scale=1.0, flag=None):
    """J: cuda f64 (m, E) tensor; det_idx: cuda int64; returns cuda f64 (n_det, m_sel).
    flag: optional cuda int32[1] tensor receiving the pivot flag (see check_flag)."""
    torch = require_cuda()
    m, ldJ = J.shape[0], J.stride(0)
    n_det = det_idx.numel()
    m_sel = rows.numel() if rows is not None else m
    out = torch.empty((n_det, m_sel), dtype=torch.float64, device=J.device)
    check(lib().ceit_inverse_model(tptr(J), ldJ, m, tptr(det_idx), n_det, tptr(rows), m_sel, float(shift),
                                   float(lmbda), int(form), float(scale), tptr(out), tptr(flag), stream_ptr()))
    return out


def inverse_model_sharded(J, det_idx_local, lmbda, shift=1.0, scale=1.0, group=None, flag=None):
    """Dual-form inverse model with the detection elements sharded over the ranks of `group`:
    partial Gram matrices are summed with one all-reduce; returns this rank's rows (n_det_local, m)."""
    torch = require_cuda()
    import torch.distributed as dist
    m, ldJ = J.shape[0], J.stride(0)
    n_loc = det_idx_local.numel()
    G = torch.empty((m, m), dtype=torch.float64, device=J.device)
    check(lib().ceit_gram_partial(tptr(J), ldJ, m, tptr(det_idx_local), n_loc, tptr(None), m, float(shift), tptr(G),
                                  stream_ptr()))
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(G, op=dist.ReduceOp.SUM, group=group)
    out = torch.empty((n_loc, m), dtype=torch.float64, device=J.device)
    if n_loc:
        check(lib().ceit_inverse_from_gram(tptr(J), ldJ, m, tptr(det_idx_local), n_loc, tptr(None), m, float(shift),
                                           float(lmbda), float(scale), tptr(G), tptr(out), tptr(flag), stream_ptr()))
    return out


def reconstruct(inv, dV, out=None):
    """inv (n_det, m) . dV (m,) or (m, F) on the device."""
    torch = require_cuda()
    n_det, m = inv.shape
    F = 1 if dV.dim() == 1 else dV.shape[1]
    if out is None:
        out = torch.empty((n_det,) if dV.dim() == 1 else (n_det, F), dtype=torch.float64, device=inv.device)
    check(lib().ceit_reconstruct(tptr(inv), tptr(dV), tptr(out), n_det, m, F, stream_ptr()))
    return out


def cuda_available():
    try:
        import torch
        return bool(torch.cuda.is_available())
    except Exception:
        return False


def mesh_model(nodes, elem, centers, radius, det_bound):
    """Device mesh model: (elem_param f64[E,9], electrode element lists, detection_index i64[n_det]) from host
    arrays (MeshObj.initialize_parameters / calc_electrode_elements / calc_detection_elements in one kernel)."""
    torch = require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device())
    nodes_d = torch.from_numpy(np.ascontiguousarray(nodes, dtype=np.float64)).to(dev)
    elem_d = torch.from_numpy(np.ascontiguousarray(elem, dtype=np.int64)).to(dev)
    N, E = nodes_d.shape[0], elem_d.shape[0]
    cen = np.ascontiguousarray(np.asarray(centers, dtype=np.float64).reshape(-1, 2))
    L = cen.shape[0]
    cen_d = torch.from_numpy(cen).to(dev)
    param = torch.empty((E, 9), dtype=torch.float64, device=dev)
    in_el = torch.empty((max(L, 1), E), dtype=torch.uint8, device=dev)
    in_det = torch.empty(E, dtype=torch.uint8, device=dev)
    check(lib().ceit_mesh_model(tptr(nodes_d), tptr(elem_d), N, E, tptr(cen_d), L, float(radius), float(det_bound),
                                tptr(param), tptr(in_el), tptr(in_det), stream_ptr()))
    pairs = torch.nonzero(in_el[:L]).cpu().numpy()            # (electrode, element), lexicographic = ascending lists
    counts = np.bincount(pairs[:, 0], minlength=L) if L else np.zeros(0, np.int64)
    lists = np.split(pairs[:, 1], np.cumsum(counts)[:-1]) if L else []
    det_idx = torch.nonzero(in_det).reshape(-1).cpu().numpy()
    return param.cpu().numpy(), [x.astype(np.int64) for x in lists], det_idx.astype(np.int64)


def center_of_position(maps, cx, cy, frac=0.9, out=None):
    """Batched DataPlotter.get_COP: maps (n_det,) or (n_det, F) cuda f64 -> (F, 4) = x, y, v means and count."""
    torch = require_cuda()
    m2 = maps if maps.dim() == 2 else maps.reshape(-1, 1)
    assert m2.is_contiguous() and m2.dtype == torch.float64 and cx.numel() == m2.shape[0] == cy.numel()
    n_det, F = m2.shape
    if out is None:
        out = torch.empty((F, 4), dtype=torch.float64, device=m2.device)
    check(lib().ceit_center_of_position(tptr(m2), n_det, F, tptr(cx), tptr(cy), float(frac), tptr(out), stream_ptr()))
    return out


def dgemm(A, B, transA=False, transB=False, alpha=1.0, beta=0.0, C=None):
    torch = require_cuda()
    M = A.shape[1] if transA else A.shape[0]
    K = A.shape[0] if transA else A.shape[1]
    N = B.shape[0] if transB else B.shape[1]
    if C is None:
        C = torch.zeros((M, N), dtype=torch.float64, device=A.device)
    check(lib().ceit_dgemm(int(transA), int(transB), M, N, K, float(alpha), tptr(A), A.stride(0), tptr(B),
                           B.stride(0), float(beta), tptr(C), C.stride(0), stream_ptr()))
    return C


def zgemm(A, B, transA=False, transB=False, alpha=1.0 + 0j, beta=0j, C=None):
    """complex128 tensors; non-conjugating transposes."""
    torch = require_cuda()
    M = A.shape[1] if transA else A.shape[0]
    K = A.shape[0] if transA else A.shape[1]
    N = B.shape[0] if transB else B.shape[1]
    if C is None:
        C = torch.zeros((M, N), dtype=torch.complex128, device=A.device)
    alpha, beta = complex(alpha), complex(beta)
    check(lib().ceit_zgemm(int(transA), int(transB), M, N, K, alpha.real, alpha.imag, tptr(A), A.stride(0), tptr(B),
                           B.stride(0), beta.real, beta.imag, tptr(C), C.stride(0), stream_ptr()))
    return C
