"""Thin loop over EFEM.calculation (host-side mirror of CEIT/Simulator.py:18-32)."""
import numpy as np

from .EFEM import EFEM
from .models.mesh import MeshObj
from .util.utilities import get_config


class Simulator:
    def __init__(self, mesh=None):
        self.mesh = MeshObj() if mesh is None else mesh
        self.config = get_config()
        self.fwd_simulator = EFEM(self.mesh)
        self.electrode_potentials = np.zeros(self.mesh.electrode_num * (self.mesh.electrode_num - 1))

    def calc_potential(self):
        L = self.mesh.electrode_num
        for i in range(L):
            _, _, electrode_potential = self.fwd_simulator.calculation(i)
            v = np.abs(np.delete(electrode_potential, i))
            self.electrode_potentials[i * (L - 1):(i + 1) * (L - 1)] = v
        return np.copy(self.electrode_potentials)
