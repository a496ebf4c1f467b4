"""Jacobian + one-shot Gauss-Newton variants (host-side mirror of CEIT/EJAC.py) on the B200 path.

The reference's double loop - for every excitation i and element j: perturb, re-assemble the dense
matrix, invert it, read the electrode amplitudes (EJAC.py:115-133) - becomes ONE call of
`ceit_jacobian_block`, which applies the exact rank-<=3 Woodbury update of a single factorised
system per (i, j).  File formats, config keys (`calc_from`/`calc_end`,
`is_first_JAC_calculation`, `variable_change_for_JAC`, `overall_origin_variable`) and the row layout
J[i(L-1)+cnt, j] are the reference's.
"""
import csv

import numpy as np

from . import _lib
from .EFEM import EFEM
from .models.mesh import MeshObj
from .util.utilities import get_config, load_npy, save_npy_async, wait_npy


class EJAC(object):
    def __init__(self, mesh=None):
        if mesh is None:
            self.mesh = MeshObj()
        else:
            self.mesh = mesh
        self.config = get_config()
        self.mode = 'n'
        self.first = self.config["is_first_JAC_calculation"]
        self.detection_bound = self.config["detection_bound"]
        self.overall_variable = self.config["overall_origin_variable"]
        self.electrode_centers = self.config["electrode_centers"]
        self.fwd_FEM = EFEM(self.mesh)
        self.electrode_num = len(self.electrode_centers)
        self.pattern_num = self.electrode_num * (self.electrode_num - 1)
        self.elem_num = self.fwd_FEM.elem_num
        self.electrode_original_potential = np.zeros((self.pattern_num))
        # JAC_matrix lives in pinned host memory (a numpy view of a pinned tensor) so that the device <-> host
        # copies of the 240 x E matrix run at PCIe speed; it is an ordinary ndarray to callers
        import torch
        self._J_pinned = torch.zeros((self.pattern_num, self.elem_num), dtype=torch.float64, pin_memory=True)
        self.JAC_matrix = self._J_pinned.numpy()
        self._inv_pinned = None
        if self.mode == 'n':
            self.fwd_FEM.reset_variable(overall_variable=self.overall_variable)
        else:
            raise Exception('No Such Mode, Please check.')
        self.initial_variable = np.copy(self.fwd_FEM.elem_variable)
        self._eng = self.fwd_FEM._engine
        self._J_dev = None
        self._J_dev_valid = False
        self._J_host_stamp = None
        self._flag_dev = None
        self._base_dev = None
        self._det_dev = None
        self.calc_origin_potential()

    # -- device helpers -----------------------------------------------------------------------------
    def _device_J(self):
        t = self._eng.torch
        if self._J_dev is None:
            self._J_dev = t.zeros((self.pattern_num, self.elem_num), dtype=t.float64, device=self._eng.dev)
        return self._J_dev

    def _upload_J(self):
        t = self._eng.torch
        J = self._device_J()
        if not self._J_is_pinned():
            # the user replaced JAC_matrix (read_JAC_np, assignment): stage it through the pinned buffer
            self._J_pinned.numpy()[...] = np.asarray(self.JAC_matrix, dtype=np.float64)
        J.copy_(self._J_pinned, non_blocking=True)
        return J

    def _J_is_pinned(self):
        Jm = self.JAC_matrix
        return (isinstance(Jm, np.ndarray) and Jm.shape == tuple(self._J_pinned.shape) and Jm.dtype == np.float64
                and Jm.ctypes.data == self._J_pinned.data_ptr())

    def calc_origin_potential(self):
        """Unperturbed electrode amplitudes for every excitation (EJAC.py:87-98)."""
        t = self._eng.torch
        if self.fwd_FEM._sync_material():
            self._eng.check_numeric()
        if self._base_dev is None:
            self._base_dev = t.zeros(self.pattern_num, dtype=t.float64, device=self._eng.dev)
        base, L, freq = self._base_dev, self.electrode_num, self.fwd_FEM.freq
        self._eng.graphed(("base", float(freq), base.data_ptr()),
                          lambda: self._eng.plan.jacobian_block(0, L, 0, 0, 0.0, freq, None, self.elem_num, base))
        self.electrode_original_potential = base.cpu().numpy()

    def JAC_calculation(self, e_from=None, e_to=None):
        """Calculate (or load) the Jacobian; saves JAC_cache.npy (EJAC.py:100-140).

        e_from/e_to (extension): restrict to an element range - the second sharding axis used by the
        multi-GPU driver (ceit_b200.sharding); the reference only has calc_from/calc_end."""
        calc_from = self.config["calc_from"]
        calc_end = self.config["calc_end"]
        variable_change = self.config["variable_change_for_JAC"]
        if self.first:
            if calc_from > 0:
                self.read_JAC_np()
            L = self.electrode_num
            i_from, i_to = max(0, calc_from), min(L, calc_end)
            e_from = 0 if e_from is None else e_from
            e_to = self.elem_num if e_to is None else e_to
            if i_to > i_from:
                self.fwd_FEM._sync_material()
                J = self._device_J()
                freq, E = self.fwd_FEM.freq, self.elem_num
                self._eng.graphed(
                    ("jac", i_from, i_to, e_from, e_to, float(variable_change), float(freq), J.data_ptr()),
                    lambda: self._eng.plan.jacobian_block(i_from, i_to, e_from, e_to, variable_change, freq, J, E))
                r0, r1 = i_from * (L - 1), i_to * (L - 1)
                wait_npy(self.config["rootdir"] + "/" + self.config["folder_name"] + '/JAC_cache.npy')   # reads the pinned matrix
                full = (i_from == 0 and i_to == L and e_from == 0 and e_to == self.elem_num)
                if self._J_is_pinned() and e_from == 0 and e_to == self.elem_num:
                    self._J_pinned[r0:r1].copy_(J[r0:r1], non_blocking=True)       # D2H into the pinned matrix
                    self._eng.torch.cuda.current_stream().synchronize()
                else:
                    self.JAC_matrix[r0:r1, e_from:e_to] = J[r0:r1, e_from:e_to].cpu().numpy()
                self._eng.check_numeric()
                # the device matrix equals the host one only when every row block came from this object
                self._J_dev_valid = bool(full and self._J_is_pinned())
                self._J_host_stamp = self._stamp_J() if self._J_dev_valid else None
            self.save_JAC_np()
            return self.JAC_matrix
        else:
            self.read_JAC_np()
            return self.JAC_matrix

    def eliminate_non_detect_JAC(self):
        """J[:, detection_index] (EJAC.py:142-152)."""
        return np.array(self.JAC_matrix[:, np.asarray(self.fwd_FEM.mesh.detection_index)])

    # -- inverse models on the device ------------------------------------------------------------------
    def _current_J(self):
        """Device copy of JAC_matrix.  After JAC_calculation() the device matrix IS the current Jacobian (the pinned
        host matrix was filled from it), so nothing is uploaded; a Jacobian the caller loaded or assigned
        (read_JAC_np, normalize_sensitivity, ...) or edited in place is staged through the pinned buffer."""
        if (self._J_dev is not None and self._J_dev_valid and self._J_is_pinned()
                and self._J_host_stamp == self._stamp_J()):
            return self._J_dev
        J = self._upload_J()
        self._J_dev_valid = True
        self._J_host_stamp = self._stamp_J()
        return J

    def _stamp_J(self):
        """Cheap fingerprint of the host matrix (four strided samples + corner sums) to notice in-place edits."""
        Jm = np.asarray(self.JAC_matrix)
        flat = Jm.reshape(-1)
        step = max(1, flat.shape[0] // 4096)
        return (Jm.shape, float(flat[::step].sum()), float(flat[:64].sum()), float(flat[-64:].sum()))

    def _inv_model(self, lmbda, shift, rows=None, scale=1.0):
        t = self._eng.torch
        J = self._current_J()
        if self._flag_dev is None:
            self._flag_dev = t.zeros(1, dtype=t.int32, device=self._eng.dev)
        flag = self._flag_dev
        det_h = np.ascontiguousarray(self.fwd_FEM.mesh.detection_index, dtype=np.int64)
        if self._det_dev is None or self._det_dev[0].shape != det_h.shape or not np.array_equal(self._det_dev[0], det_h):
            self._det_dev = (det_h.copy(), t.from_numpy(det_h).to(self._eng.dev))
        det = self._det_dev[1]
        if rows is not None:
            rows_d = t.tensor(rows, dtype=t.int64, device=self._eng.dev)
            inv = _lib.inverse_model(J, det, lmbda, rows=rows_d, shift=shift, scale=scale, flag=flag)
        else:
            # the result lives in the graph's memory pool: valid until the same model is computed again
            inv = self._eng.graphed(
                ("inv", float(lmbda), float(shift), float(scale), J.data_ptr(), det.data_ptr()),
                lambda: _lib.inverse_model(J, det, lmbda, shift=shift, scale=scale, flag=flag))
        _lib.check_flag(flag)          # np.linalg.inv raises on a singular matrix (EJAC.py:165,241): never cache garbage
        return inv

    def _apply(self, inv, delta_V):
        t = self._eng.torch
        dV = t.from_numpy(np.ascontiguousarray(delta_V, dtype=np.float64)).to(self._eng.dev)
        return _lib.reconstruct(inv, dV).cpu().numpy()

    def eit_solve(self, detect_potential, lmbda=295):
        """EJAC.py:154-170."""
        delta_V = detect_potential - np.copy(self.electrode_original_potential)
        return self._apply(self._inv_model(lmbda, 1.0), delta_V)

    def eit_solve_delta_V(self, delta_V, lmbda=295):
        """EJAC.py:172-182 (the reference does NOT subtract 1 here; kept as is)."""
        return self._apply(self._inv_model(lmbda, 0.0), delta_V)

    def eit_solve_4electrodes(self, detect_potential, lmbda=90):
        return self.eit_solve_on_some_electrodes([2, 6, 10, 14], detect_potential, lmbda, mode="direct")

    def eit_solve_4electrodes_delta_V(self, delta_V, lmbda=90):
        return self.eit_solve_on_some_electrodes([2, 6, 10, 14], delta_V, lmbda, mode="diff")

    def eit_solve_8electrodes(self, detect_potential, lmbda=265):
        return self.eit_solve_on_some_electrodes([0, 2, 4, 6, 8, 10, 12, 14], detect_potential, lmbda, mode="direct")

    def eit_solve_8electrodes_delta_V(self, delta_V, lmbda=265):
        return self.eit_solve_on_some_electrodes([0, 2, 4, 6, 8, 10, 12, 14], delta_V, lmbda, mode="diff")

    def eit_solve_on_some_electrodes(self, electrode_list, potential, lmbda, mode="diff"):
        """Row-subset inverse model (EJAC.py:184-229), result scaled by 1e-4 (EJAC.py:228)."""
        assert mode == "diff" or mode == "direct", "Please enter the right solver mode"
        slice_list = []
        for i in electrode_list:
            for j in electrode_list:
                if j < i:
                    slice_list.append(i * (self.electrode_num - 1) + j)
                if j > i:
                    slice_list.append(i * (self.electrode_num - 1) + j - 1)
        if mode == "direct":
            delta_V = potential - np.copy(self.electrode_original_potential)
            delta_V = np.copy(delta_V[slice_list])
        else:
            delta_V = potential
        return self._apply(self._inv_model(lmbda, 1.0, rows=slice_list, scale=1e-4), delta_V)

    def save_inv_matrix(self, lmbda=203):
        """EJAC.py:231-243."""
        inv = self._inv_model(lmbda, 1.0)
        t = self._eng.torch
        path = self.config["rootdir"] + "/" + self.config["folder_name"] + '/inv_mat.npy'
        wait_npy(path)                                  # the previous write reads the pinned buffer
        if self._inv_pinned is None or tuple(self._inv_pinned.shape) != tuple(inv.shape):
            self._inv_pinned = t.empty(tuple(inv.shape), dtype=t.float64, pin_memory=True)
        self._inv_pinned.copy_(inv, non_blocking=True)
        t.cuda.current_stream().synchronize()
        save_npy_async(path, self._inv_pinned.numpy())  # write-behind (joined before any reader, see utilities)

    def read_inv_matrix(self):
        return load_npy(self.config["rootdir"] + "/" + self.config["folder_name"] + '/inv_mat.npy')

    def eit_solve_direct(self, detect_potential):
        """EJAC.py:252-259."""
        t = self._eng.torch
        JAC_p = t.from_numpy(self.read_inv_matrix()).to(self._eng.dev)
        delta_V = detect_potential - np.copy(self.electrode_original_potential)
        return self._apply(JAC_p, delta_V)

    # -- persistence (EJAC.py:261-294) ------------------------------------------------------------------
    def save_JAC_np(self):
        path = self.config["rootdir"] + "/" + self.config["folder_name"] + '/JAC_cache.npy'
        if self._J_is_pinned():
            save_npy_async(path, self.JAC_matrix)       # write-behind: overlaps the inverse model on the GPU
        else:
            wait_npy(path)
            save_npy_async(path, np.array(self.JAC_matrix, dtype=np.float64))
            wait_npy(path)

    def read_JAC_np(self):
        self.JAC_matrix = load_npy(self.config["rootdir"] + "/" + self.config["folder_name"] + '/JAC_cache.npy')

    def save_JAC_2file(self):
        with open(self.config["rootdir"] + "/" + self.config["folder_name"] + '/jac_cache.csv', "w",
                  newline='') as csvfile:
            writer = csv.writer(csvfile, delimiter=',')
            for row in self.JAC_matrix:
                writer.writerow(row)

    def read_JAC_2file(self, mode='normal'):
        wait_npy()                                     # a pending cache write reads JAC_matrix
        if mode == 'normal':
            with open(self.config["rootdir"] + "/" + self.config["folder_name"] + '/jac_cache.csv',
                      newline='') as csvfile:
                for i, line in enumerate(csv.reader(csvfile, delimiter=',')):
                    self.JAC_matrix[i] = line
        else:
            raise Exception('Mode do not exist.')

    def get_sensitivity_list(self):
        """EJAC.py:303-308."""
        return np.sum(self.JAC_matrix, axis=0)

    def normalize_sensitivity(self):
        """DEPRECATED in the reference (EJAC.py:310-318); kept for interface parity."""
        sensitivity = self.get_sensitivity_list()
        self.JAC_matrix = self.JAC_matrix / sensitivity.T * np.mean(sensitivity)

    def show_JAC(self):
        raise NotImplementedError("plotting is outside the scope of ceit_b200")

    plot_sensitivity = plot_map_in_detection_range = show_JAC
