"""Realtime reconstructor (host-side mirror of CEIT/Solver.py) on the B200 path.

`Solver(lmbda).solve(delta_V)` keeps the reference's behaviour: J is read from JAC_cache.npy, the
inverse model inv(Jt^T Jt + lmbda^2 I) Jt^T with Jt = J[:, detection_index] - 1 is built once
(Solver.py:110-124) and cached in inv_mat.npy, and each frame is one product inv_mat . delta_V
(Solver.py:95-108).  Both steps run on the device; `inv_mat` stays resident in HBM between frames.
The cache-existence test is CWD-relative exactly like the reference's (Solver.py:44) while load and
save use `rootdir` (Solver.py:92-93,123-124).
"""
import os

import numpy as np

from . import _lib
from .models.mesh import MeshObj
from .readmesh import read_mesh_from_csv
from .util.utilities import get_config, load_npy, save_npy


class Solver(object):
    def __init__(self, lmbda=289):
        print("Please make sure info in config.json is all correct, HERE WE GO!")
        self.config = get_config()
        mesh_obj, electrode_num, electrode_center_list, electrode_radius = read_mesh_from_csv().return_mesh()
        self.mesh = MeshObj(mesh_obj, electrode_num, electrode_center_list, electrode_radius)
        self._torch = _lib.require_cuda()
        self._dev = self._torch.device("cuda", self._torch.cuda.current_device())
        self._inv_dev = None
        self.read_JAC()
        self.elem_param = np.zeros((np.shape(self.mesh.elem)[0], 9))
        self.inv_mat = None
        if os.path.exists(self.config["folder_name"] + '/inv_mat.npy'):
            self.read_inv_matrix()
            assert self.inv_mat.shape[0] == self.mesh.detection_index.shape[
                0], "inverse matrix file inv_mat.npy dimension not equal with detection area,\n please run " \
                    "reinitialize_solver() function from CEIT.Solver to generate a new one or delete the file. "
            assert self.inv_mat.shape[1] == self.mesh.electrode_num * (
                self.mesh.electrode_num - 1), "Inverse matrix pattern dimension does not match electrode number, " \
                                              "please check your JAC matrix, aborting ... "
        else:
            self.read_JAC()
            self.get_inv_matrix(lmbda)

    def return_mesh_info(self):
        return self.mesh.point_x, self.mesh.point_y, self.mesh.elem, self.mesh.detection_elem

    def _path(self, name):
        return self.config["rootdir"] + "/" + self.config["folder_name"] + "/" + name

    def read_JAC(self):
        assert os.path.exists(self._path('JAC_cache.npy')), "The JAC matrix is not generated"
        self.JAC_mat = load_npy(self._path('JAC_cache.npy'))

    def eliminate_non_detect_JAC(self):
        """Solver.py:76-86."""
        return np.array(self.JAC_mat[:, np.asarray(self.mesh.detection_index)])

    def read_inv_matrix(self):
        self.inv_mat = load_npy(self._path('inv_mat.npy'))
        self._inv_dev = None

    def _inv_on_device(self):
        t = self._torch
        if self._inv_dev is None:
            self._inv_dev = t.from_numpy(np.ascontiguousarray(self.inv_mat, dtype=np.float64)).to(self._dev)
        return self._inv_dev

    def solve(self, delta_V):
        """delta_V: (m,) or (m, F) host array -> density prediction (n_det,) or (n_det, F)."""
        t = self._torch
        dV = t.from_numpy(np.ascontiguousarray(delta_V, dtype=np.float64)).to(self._dev, non_blocking=True)
        return _lib.reconstruct(self._inv_on_device(), dV).cpu().numpy()

    def solve_device(self, delta_V_dev, out=None):
        """Same product for frames already resident in HBM (cuda f64 tensor (m,) or (m, F))."""
        return _lib.reconstruct(self._inv_on_device(), delta_V_dev, out=out)

    def get_inv_matrix(self, lmbda=289, form=0):
        """Build + save the inverse model (Solver.py:110-124). form: 0 auto, 1 primal, 2 dual."""
        t = self._torch
        self.read_JAC()
        J = t.from_numpy(np.ascontiguousarray(self.JAC_mat, dtype=np.float64)).to(self._dev)
        det = t.from_numpy(np.ascontiguousarray(self.mesh.detection_index, dtype=np.int64)).to(self._dev)
        flag = t.zeros(1, dtype=t.int32, device=self._dev)
        self._inv_dev = _lib.inverse_model(J, det, lmbda, shift=1.0, form=form, flag=flag)
        _lib.check_flag(flag)          # np.linalg.inv raises on a singular matrix (Solver.py:121): nothing is cached
        self.inv_mat = self._inv_dev.cpu().numpy()
        save_npy(self._path('inv_mat.npy'), self.inv_mat)

    def adaptive_solver(self, delta_V):
        """NOT COMPLETED in the reference either (Solver.py:126-130)."""
        pass

    def plot_map_in_detection_range(self, ax, param, vmax=None, vmin=None):
        raise NotImplementedError("plotting is outside the scope of ceit_b200")


def reinitialize_solver(lmbda=0):
    """Delete inv_mat.npy and regenerate it (Solver.py:150-167)."""
    assert (lmbda > 0), "Please enter a positive lmbda parameter"
    config = get_config()
    path = config["rootdir"] + "/" + config["folder_name"] + "/" + 'inv_mat.npy'
    if os.path.exists(path):
        os.remove(path)
    return Solver(lmbda)
