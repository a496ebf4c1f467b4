"""Device-side state shared by EFEM / EJAC for one mesh: the native plan, the uploaded material
vectors and the factorisation cache.  PyTorch tensors are plain device buffers here."""
import numpy as np

from . import _lib


class Engine:
    """One per MeshObj (cached on the mesh as `_ceit_engine`)."""

    def __init__(self, mesh, target_block=0):
        torch = _lib.require_cuda()
        self.torch = torch
        self.dev = torch.device("cuda", torch.cuda.current_device())
        self.L = mesh.electrode_num
        electrodes = [mesh.electrode_mesh[i] for i in range(self.L)]     # dict order == index order
        self.plan = _lib.Plan(mesh.nodes, mesh.elem, electrodes, target_block=target_block)
        self.N, self.E = self.plan.N, self.plan.E
        self._perm_host = None
        self._var_host = None
        self._omega = None
        self._factored = None        # None | False (real) | True (complex)
        self.perm_d = torch.empty(self.E, dtype=torch.float64, device=self.dev)
        self.var_d = torch.empty(self.E, dtype=torch.float64, device=self.dev)
        # CUDA graphs of library-call sequences on static buffers, keyed by their frozen scalar arguments
        self.use_graphs = True
        self._graphs = {}
        self._graph_gen = self.plan.generation()

    @staticmethod
    def for_mesh(mesh):
        eng = getattr(mesh, "_ceit_engine", None)
        if eng is None or eng.N != np.shape(mesh.nodes)[0] or eng.E != np.shape(mesh.elem)[0]:
            eng = Engine(mesh)
            mesh._ceit_engine = eng
        return eng

    MAX_GRAPHS = 16

    def graphed(self, key, fn):
        """Run `fn` (library calls on buffers owned by this engine / its users, all arguments determined by
        `key`).  The first call per key runs eagerly (workspace allocation), the second is captured into a
        CUDA graph, later ones replay it: no launch gaps between the ~100 short kernels of a step."""
        if not self.use_graphs or _lib._options.get("stage_timers", 0):
            return fn()
        # A captured graph holds the plan's workspace pointers and kernel choices of the allocation generation it
        # was captured in; a real <-> complex switch or a larger excitation chunk re-allocates them (the library
        # bumps the generation), after which every older graph is stale and is dropped.
        gen = self.plan.generation()
        if gen != self._graph_gen:
            self._graphs.clear()
            self._graph_gen = gen
        ent = self._graphs.get(key)
        if ent is None:
            out = fn()                               # eager: may allocate (never inside a capture)
            gen2 = self.plan.generation()
            if gen2 != gen:
                self._graphs.clear()
                self._graph_gen = gen2
            if len(self._graphs) >= self.MAX_GRAPHS:
                self._graphs.pop(next(iter(self._graphs)))
            self._graphs[key] = "warm"
            return out
        if ent == "warm":
            try:
                ent = self._graphs[key] = _lib.StepGraph(fn)
            except _lib.CeitError:
                # the library refused to allocate inside the capture: stay eager for this key
                self._graphs.pop(key, None)
                return fn()
        return ent.replay()

    def set_material(self, perm, var, omega):
        """Upload perm / var if they changed; (re)assemble + factor when needed.
        Returns True when a new factorisation was computed."""
        torch = self.torch
        perm = np.ascontiguousarray(perm, dtype=np.float64)
        var = np.ascontiguousarray(var, dtype=np.float64)
        same = (self._factored is not None and self._omega == omega and
                np.array_equal(perm, self._perm_host) and np.array_equal(var, self._var_host))
        if same:
            return False
        self.perm_d.copy_(torch.from_numpy(perm))
        self.var_d.copy_(torch.from_numpy(var))
        self._perm_host, self._var_host, self._omega = perm.copy(), var.copy(), omega
        is_complex = bool(np.any(var != 0.0))

        def factor():
            self.plan.assemble(self.perm_d, self.var_d, omega)
            self.plan.factor(is_complex)
        self.graphed(("factor", float(omega), is_complex), factor)
        self._factored = is_complex
        return True

    def csr_values(self):
        """complex128[nnz] of the last assembled K (downloads)."""
        torch = self.torch
        vals = torch.empty((self.plan.nnz, 2), dtype=torch.float64, device=self.dev)
        self.plan.csr_values(vals)
        v = vals.cpu().numpy()
        return v[:, 0] + 1j * v[:, 1]

    def check_numeric(self):
        if self.plan.numeric_flag() != 0:
            raise Exception("Matrix is not positive definite: check mesh connectivity / electrode layout")
