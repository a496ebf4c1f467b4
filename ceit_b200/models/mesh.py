"""Mesh model (host-side mirror of CEIT/models/mesh.py): the data contract the kernels consume.

Attributes: nodes f64[N,2] (metres), elem int64[E,3], elem_perm f64[E],
elem_param f64[E,9] = [A, b0,b1,b2, c0,c1,c2, xbar, ybar], electrode_mesh {i: [element ids]},
detection_index int64[n_det] ascending, detection_elem int64[n_det,3].
The per-element Python loops with 3x3 np.linalg.inv of the reference (mesh.py:41-75) are closed-form
expressions here (agreement ~1e-16 relative).  With a CUDA device the three per-element passes - elem_param, the
electrode element sets and the detection set - run as ONE kernel (`ceit_mesh_model`); the element sets it produces are
bit-identical to the host expressions below, which remain for hosts without a GPU (mesh preparation such as
`init_mesh`, the CPU test-suite) - the model is input preparation, not the compute path.
"""
import numpy as np

from .. import _lib

from ..readmesh import read_mesh_from_csv
from ..util.utilities import get_config


class MeshObj(object):
    def __init__(self, mesh_obj=None, electrode_num=None, electrode_center_list=None, electrode_radius=None):
        self.electrode_mesh = dict()
        if mesh_obj is None or electrode_num is None or electrode_center_list is None or electrode_radius is None:
            mesh_obj, electrode_num, electrode_center_list, electrode_radius = read_mesh_from_csv().return_mesh()
        self.config = get_config()
        self.mesh_obj = mesh_obj
        self.electrode_num = electrode_num
        self.electrode_center_list = electrode_center_list
        self.electrode_radius = electrode_radius
        self.nodes = self.mesh_obj["node"]
        self.point_x = self.nodes[:, 0]
        self.point_y = self.nodes[:, 1]
        self.elem = self.mesh_obj["element"]
        self.elem_perm = self.mesh_obj["perm"]
        self.detection_index = np.zeros((len(self.elem)))
        self.detection_elem = np.copy(self.elem)
        self.elem_param = np.zeros((np.shape(self.elem)[0], 9))
        for i in range(self.electrode_num):
            self.electrode_mesh[i] = list()
        self.on_device = False
        if _lib.cuda_available() and self.electrode_num > 0:
            self._device_model()
        else:
            self.initialize_parameters()
            self.calc_detection_elements()

    def _device_model(self):
        """elem_param, electrode element sets and detection set in one kernel launch (ceit_mesh_model)."""
        param, lists, det_idx = _lib.mesh_model(self.nodes, self.elem, self.electrode_center_list,
                                                self.electrode_radius, self.config["detection_bound"])
        self.elem_param = param
        for i in range(self.electrode_num):
            if len(lists[i]) == 0:
                raise Exception("No element is selected, please check the input")       # mesh.py:111-113
            self.electrode_mesh[i] = [int(e) for e in lists[i]]
        self.detection_index = det_idx
        self.detection_elem = np.array(self.elem[det_idx])
        self.on_device = True

    def initialize_parameters(self):
        """elem_param + electrode element sets (mesh.py:41-75)."""
        x = self.nodes[self.elem, 0]
        y = self.nodes[self.elem, 1]
        det = (x[:, 0] * (y[:, 1] - y[:, 2]) - y[:, 0] * (x[:, 1] - x[:, 2])
               + (x[:, 1] * y[:, 2] - x[:, 2] * y[:, 1]))
        self.elem_param[:, 0] = np.abs(det / 2)
        self.elem_param[:, 1] = (y[:, 1] - y[:, 2]) / det
        self.elem_param[:, 2] = (y[:, 2] - y[:, 0]) / det
        self.elem_param[:, 3] = (y[:, 0] - y[:, 1]) / det
        self.elem_param[:, 4] = (x[:, 2] - x[:, 1]) / det
        self.elem_param[:, 5] = (x[:, 0] - x[:, 2]) / det
        self.elem_param[:, 6] = (x[:, 1] - x[:, 0]) / det
        self.elem_param[:, 7] = x.mean(1)
        self.elem_param[:, 8] = y.mean(1)
        for i in range(self.electrode_num):
            self.calc_electrode_elements(i, self.electrode_center_list[i], self.electrode_radius)

    def scale_mesh(self, scale_factor):
        """new_length = old_length * scale_factor (mesh.py:77-88)."""
        self.nodes *= scale_factor
        self.electrode_radius *= scale_factor
        self.electrode_center_list = list(np.array(self.electrode_center_list) * scale_factor)

    def calc_electrode_elements(self, electrode_number, center, radius):
        """Elements whose centroid lies in the closed square around `center` (mesh.py:90-113)."""
        if electrode_number >= self.electrode_num:
            raise Exception("the input number exceeded electrode numbers")
        center_x, center_y = center
        px, py = self.elem_param[:, 7], self.elem_param[:, 8]
        sel = np.nonzero(((center_x + radius) >= px) & (px >= (center_x - radius)) &
                         ((center_y + radius) >= py) & (py >= (center_y - radius)))[0]
        self.electrode_mesh[electrode_number] += [int(i) for i in sel]
        if len(sel) == 0:
            raise Exception("No element is selected, please check the input")

    def return_mesh(self):
        return self.mesh_obj

    def return_electrode_info(self):
        return self.electrode_center_list, self.electrode_radius

    def calc_detection_elements(self):
        """Elements whose centroid is strictly inside the detection square (mesh.py:127-155)."""
        n, e = self.nodes, self.elem
        x_val = (n[e[:, 0], 0] + n[e[:, 1], 0] + n[e[:, 2], 0]) / 3
        y_val = (n[e[:, 0], 1] + n[e[:, 1], 1] + n[e[:, 2], 1]) / 3
        b = self.config["detection_bound"]
        idx = np.nonzero((np.abs(x_val) < b) & (np.abs(y_val) < b))[0]
        self.detection_index = np.array(idx)
        self.detection_elem = np.array(self.elem[idx])

    def delete_outside_detect(self, list_c):
        """Keep the detection-area entries of an element-wise list (mesh.py:157-180)."""
        list_c = np.array(list_c)
        if list_c.ndim >= 1:
            return np.array(list_c[self.detection_index], dtype=np.float64)
        raise Exception("Transfer Shape Not Correct")

    def change_conductivity(self, element_list, resistance_list):
        """mesh.py:182-196."""
        if len(element_list) == len(resistance_list):
            for i, ele_num in enumerate(element_list):
                if ele_num > len(self.elem):
                    raise Exception("Element number exceeds limit")
                self.elem_perm[ele_num] = resistance_list[i]
        else:
            raise Exception('The length of element doesn\'t match the length of variable')

    def get_perimeter(self):
        """Convex hull of the nodes as a CCW chain of node indices starting at the lowest (then
        leftmost) node (mesh.py:198-251).  Andrew's monotone chain instead of the reference's
        hand-written quicksort + Graham scan."""
        pts = self.nodes
        order = np.lexsort((pts[:, 1], pts[:, 0]))

        def cross(o, a, b):
            return (pts[a, 0] - pts[o, 0]) * (pts[b, 1] - pts[o, 1]) - (pts[a, 1] - pts[o, 1]) * (pts[b, 0] - pts[o, 0])

        lower, upper = [], []
        for i in order:
            while len(lower) >= 2 and cross(lower[-2], lower[-1], i) <= 0:
                lower.pop()
            lower.append(int(i))
        for i in order[::-1]:
            while len(upper) >= 2 and cross(upper[-2], upper[-1], i) <= 0:
                upper.pop()
            upper.append(int(i))
        hull = lower[:-1] + upper[:-1]
        start = min(range(len(hull)), key=lambda k: (pts[hull[k], 1], pts[hull[k], 0]))
        return np.array(hull[start:] + hull[:start], dtype=int)
