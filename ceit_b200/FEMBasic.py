"""Base class of the forward models (host-side mirror of CEIT/FEMBasic.py).

Same public state and methods as the reference; the arithmetic of `calculation()` runs on the GPU
through the C ABI (see EFEM.my_solver).  Differences forced by scale:
`K_sparse` (dense N x N complex in the reference, FEMBasic.py:61) is materialised lazily on
attribute access instead of being re-allocated on every solve, and `K_node_num_list`
(the swap bookkeeping, FEMBasic.py:62) is the identity because nothing is swapped any more.
"""
from abc import ABCMeta, abstractmethod

import numpy as np

from .models.mesh import MeshObj
from .util.utilities import get_config


class FEMBasic(metaclass=ABCMeta):
    def __init__(self, mesh=None):
        self.config = get_config()
        if mesh is None:
            self.mesh = MeshObj()
        else:
            self.mesh = mesh
        self.node_num = np.shape(self.mesh.nodes)[0]
        if np.shape(self.mesh.nodes)[1] != 2:
            raise Exception("2D Node coordinates incorrect")
        self.node_num_bound = 0
        self.node_num_f = 0
        self.node_num_ground = 0
        if np.shape(self.mesh.elem)[1] != 3:
            raise Exception("2D Elements dimension incorrect")
        self.elem_num = np.shape(self.mesh.elem)[0]
        self.perm = 1 / self.config["resistance"]
        self.mesh.elem_perm = self.mesh.elem_perm * self.perm      # in place on the shared mesh (FEMBasic.py:58)
        self.elem_variable = np.zeros(np.shape(self.mesh.elem_perm))
        self.freq = self.config["signal_frequency"] * 2 * np.pi
        self.node_potential = np.zeros((self.node_num), dtype=np.complex128)
        self.element_potential = np.zeros((self.elem_num), dtype=np.complex128)
        self.electrode_potential = np.zeros((self.mesh.electrode_num), dtype=np.complex128)

    # -- dense K on demand -------------------------------------------------------------------------
    @property
    def K_sparse(self):
        if self.node_num > 20000:
            raise MemoryError("K_sparse as a dense N x N complex matrix needs %.1f GB; use csr_K()"
                              % (16e-9 * self.node_num ** 2))
        rowptr, colidx, vals = self.csr_K()
        K = np.zeros((self.node_num, self.node_num), dtype=np.complex128)
        rows = np.repeat(np.arange(self.node_num), np.diff(rowptr))
        K[rows, colidx] = vals
        return K

    @property
    def K_node_num_list(self):
        return [x for x in range(self.node_num)]

    def calculation(self, electrode_input):
        """Forward calculation -> (node_potential, element_potential, electrode_potential)
        (FEMBasic.py:67-86)."""
        self.calc_init()
        self.my_solver(electrode_input)
        self.calculate_element_potential()
        self.calc_electrode_potential()
        return self.node_potential, self.element_potential, self.electrode_potential

    @abstractmethod
    def my_solver(self, electrode_input):
        pass

    def calc_init(self):
        """FEMBasic.py:105-114 (the dense K is no longer re-zeroed: it is never formed)."""
        self.node_potential = np.zeros((self.node_num), dtype=np.complex128)
        self.element_potential = np.zeros((self.elem_num), dtype=np.complex128)
        self.electrode_potential = np.zeros((self.mesh.electrode_num), dtype=np.complex128)

    @abstractmethod
    def calculate_element_potential(self):
        pass

    @abstractmethod
    def calc_electrode_potential(self):
        pass

    # -- variable editing (host state; uploaded lazily before the next solve) -----------------------
    def change_variable_elementwise(self, element_list, variable_list):
        """FEMBasic.py:163-179."""
        if len(element_list) == len(variable_list):
            for i, ele_num in enumerate(element_list):
                if ele_num > self.elem_num:
                    raise Exception("Element number exceeds limit")
                self.elem_variable[ele_num] = variable_list[i]
        else:
            raise Exception('The length of element doesn\'t match the length of variable')

    def _select_geometry(self, center, radius, shape):
        cx, cy = center
        px, py = self.mesh.elem_param[:, 7], self.mesh.elem_param[:, 8]
        if shape == "square":
            sel = ((cx + radius) >= px) & (px >= (cx - radius)) & ((cy + radius) >= py) & (py >= (cy - radius))
            if not np.any(sel):
                raise Exception("No element is selected, please check the input")
            return sel
        if shape == "circle":
            return np.sqrt((cx - px) ** 2 + (cy - py) ** 2) <= radius
        raise Exception("No such shape, please check the input")

    def change_variable_geometry(self, center, radius, value, shape):
        """FEMBasic.py:181-212."""
        self.elem_variable[self._select_geometry(center, radius, shape)] = value

    def change_add_variable_geometry(self, center, radius, value, shape):
        """FEMBasic.py:214-246."""
        self.elem_variable[self._select_geometry(center, radius, shape)] += value

    def reset_variable(self, overall_variable=0):
        """FEMBasic.py:248-255."""
        self.elem_variable = np.zeros(np.shape(self.mesh.elem_perm)) + overall_variable

    def change_conductivity(self, element_list, resistance_list):
        """FEMBasic.py:257-265."""
        self.mesh.change_conductivity(element_list, resistance_list)

    # -- plotting: out of scope ------------------------------------------------------
    def plot_map(self, ax, param):
        raise NotImplementedError("plotting is outside the scope of ceit_b200 (matplotlib layer of the reference)")

    plot_variable_map = plot_potential_map = plot_map
