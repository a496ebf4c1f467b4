"""Centre-of-position pipeline over recorded data (host-side mirror of CEIT/datahandler/DataHandler.py).

`get_COP(data)` keeps the reference contract (DataHandler.py:28-56): reconstruct the map of one recording and
return the mean centroid and mean value of the detection elements above 90 % of the map's range.  The whole
dataset goes through `get_COP_batch`, ONE batched reconstruction (`ceit_reconstruct`, frames resident in HBM) and
ONE reduction kernel (`ceit_center_of_position`) instead of a Python loop of solves.  A "recording" here is anything
with a `delta_V` attribute (m,), as the reference's `Data` objects have after `calc_delta_V`; the lab-file classes
(`Data`, `CalibData`, `DataSet`, csv readers) and plotting are outside the scope of ceit_b200.
"""
import numpy as np

from .. import _lib
from ..Solver import Solver, reinitialize_solver
from ..util.utilities import get_config

THRESHOLD_PERCENTAGE = 0.9


class DataPlotter:
    def __init__(self, solver=None):
        if solver is None:
            self.solver = reinitialize_solver(get_config()["regularization_coeff"])
        else:
            self.solver = solver
        self._centroids = None

    def _centroids_dev(self):
        """(cx, cy) of the detection elements on the device (elem_param[:, 7:9], mesh.py:70-71)."""
        if self._centroids is None:
            torch = _lib.require_cuda()
            mesh = self.solver.mesh
            par = np.asarray(mesh.elem_param)[np.asarray(mesh.detection_index)]
            dev = torch.device("cuda", torch.cuda.current_device())
            self._centroids = (torch.from_numpy(np.ascontiguousarray(par[:, 7])).to(dev),
                               torch.from_numpy(np.ascontiguousarray(par[:, 8])).to(dev))
        return self._centroids

    def get_COP_batch(self, delta_V):
        """delta_V: (m, F) host array, one recording per column -> (x_mean, y_mean, v_mean, count), each (F,).
        Frames where no element exceeds the threshold (a constant map) have count 0 and NaN means."""
        torch = _lib.require_cuda()
        dV = np.ascontiguousarray(delta_V, dtype=np.float64)
        if dV.ndim != 2:
            raise Exception("get_COP_batch expects an (m, F) array")
        dev = torch.device("cuda", torch.cuda.current_device())
        maps = self.solver.solve_device(torch.from_numpy(dV).to(dev))
        cx, cy = self._centroids_dev()
        res = _lib.center_of_position(maps, cx, cy, THRESHOLD_PERCENTAGE).cpu().numpy()
        return res[:, 0], res[:, 1], res[:, 2], res[:, 3]

    def get_COP(self, data):
        """x, y of the centre and the mean value above the threshold (DataHandler.py:28-56)."""
        if data.delta_V is None:
            raise Exception("WRONG!")
        x, y, v, count = self.get_COP_batch(np.asarray(data.delta_V, dtype=np.float64).reshape(-1, 1))
        if count[0] == 0:
            raise ZeroDivisionError("float division by zero")
        return float(x[0]), float(y[0]), float(v[0])

    def draw_COP(self, dataset, calibration, ax, title="COP"):
        raise NotImplementedError("plotting is outside the scope of ceit_b200 (the numbers: cop_table)")

    def draw_one_situation(self, data, calibration, ax, vmax=None, vmin=None, title_prefix=""):
        raise NotImplementedError("plotting is outside the scope of ceit_b200")


def get_corr_from_filename(filename):
    parts = filename.split('.')[0].split('_')
    return int(parts[3]), int(parts[4]), int(parts[5])


def get_height_idx_from_filename(filename):
    return int(filename.split('.')[0].split('_')[-1])
