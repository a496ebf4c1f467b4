"""EFEM forward solver (host-side mirror of CEIT/EFEM.py) on the B200 path.

`calculation(i)` keeps the reference's contract (FEMBasic.py:67-86): Dirichlet potential 1 on the
nodes of excitation electrode i, complex linear-triangle system, returns |phi| per node, the mean of
the 3 node amplitudes per element and the mean over each electrode's elements.  Instead of
re-assembling a dense N x N matrix and inverting it per call (EFEM.py:47-69, 98-126), the material
state (elem_perm, elem_variable) is assembled and factorised ONCE on the device and every excitation
re-uses that factorisation through a Schur complement on the electrode nodes.
"""
import numpy as np

from ._engine import Engine
from .FEMBasic import FEMBasic


class EFEM(FEMBasic):
    def __init__(self, mesh=None):
        super().__init__(mesh)
        dev = self.config["device"]
        if dev == "cpu":
            raise Exception('ceit_b200 has no CPU path: set "device": "gpu" in config.json '
                            '(the reference numpy path is only kept as the test oracle)')
        if dev != "gpu":
            raise Exception('Please make sure you specified device inside the config file to "cpu" or "gpu"')
        self._engine = Engine.for_mesh(self.mesh)
        t = self._engine.torch
        self._d_node = t.empty(self.node_num, dtype=t.float64, device=self._engine.dev)
        self._d_elem = t.empty(self.elem_num, dtype=t.float64, device=self._engine.dev)
        self._d_elec = t.empty(self.mesh.electrode_num, dtype=t.float64, device=self._engine.dev)

    def _sync_material(self):
        return self._engine.set_material(self.mesh.elem_perm, self.elem_variable, self.freq)

    def invalidate(self):
        """Forget the cached factorisation: the next solve re-assembles and re-factorises."""
        self._engine._factored = None

    def my_solver(self, electrode_input):
        """Assemble + factor (cached per material state), then one Schur solve (EFEM.py:37-45)."""
        if not (0 <= electrode_input < self.mesh.electrode_num):
            raise IndexError("list index out of range")      # electrode_list[electrode_input], EFEM.py:81
        if self._sync_material():
            self._engine.check_numeric()
        self._engine.graphed(("fwd", int(electrode_input), self._d_node.data_ptr()),
                             lambda: self._engine.plan.forward(electrode_input, self._d_node, self._d_elem,
                                                               self._d_elec))
        nb = len(np.unique(self.mesh.elem[np.asarray(self.mesh.electrode_mesh[electrode_input])].reshape(-1)))
        self.node_num_bound = nb
        self.node_num_f = self.node_num - nb
        self.node_potential = self._d_node.cpu().numpy()      # |phi| per node, original order (EFEM.py:44,128-134)

    def calculate_element_potential(self):
        """FEMBasic.py:118-127 - computed by elem_mean_kernel in the same call as the solve."""
        self.element_potential = self._d_elem.cpu().numpy().astype(np.complex128)

    def calc_electrode_potential(self):
        """FEMBasic.py:129-137 - computed by baseline_kernel in the same call as the solve."""
        self.electrode_potential = self._d_elec.cpu().numpy().astype(np.complex128)

    def construct_sparse_matrix(self):
        """Assemble on the device (EFEM.py:47-69); the values are available through csr_K()."""
        self._sync_material()

    def csr_K(self):
        """(rowptr, colidx, complex values) of K in original numbering; equals
        scipy.sparse.csr_matrix(reference K_sparse) with sorted indices."""
        self._sync_material()
        tb = self._engine.plan.tables()
        return tb["rowptr"], tb["colidx"], self._engine.csr_values()
