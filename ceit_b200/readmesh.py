"""Mesh ingestion and the mesh cache (host-side mirror of CEIT/readmesh.py).

File formats are the reference's (they are the drop-in contract):
  <rootdir>/<folder_name>/EITMesh.fem          OptiStruct deck, Nastran short-field GRID / CTRIA3
  cache_Mesh_nodes.pkl / cache_Mesh_elements.pkl   pickled list of ndarray(2,) / list of [int]*3
  Mesh_Cache_Node.csv / Mesh_Cache_Element.csv     one row per node / element
The O(N^2 E) duplicate-node scan of the reference (readmesh.py:92-113) is replaced by a hash map
with the same result (first occurrence keeps its place, later duplicates are merged into it).
"""
import csv
import os
import re

import numpy as np

from .util.utilities import get_config, read_parameter, save_parameter


def string_to_float(input_num):
    """Nastran short-field float: '1.5-3' means 1.5e-3 (readmesh.py:19-31)."""
    if re.match(r"[0-9\.]+\-[0-9]+", input_num):
        num, power = input_num.split("-")
        return float(num) * 10 ** (-int(power))
    if re.match(r"\-[0-9\.]+\-[0-9]+", input_num):
        _, num, power = input_num.split("-")
        return -float(num) * 10 ** (-int(power))
    return float(input_num)


def parse_fem(path):
    """GRID x,y in columns 24:32 / 32:40; CTRIA3 node ids in 24:32, 32:40, 40:48 (readmesh.py:50-61)."""
    nodes, elements = [], []
    with open(path, newline="") as file:
        for line in file:
            if line.find("$") != -1:
                continue
            if line.find("GRID") != -1:
                nodes.append([string_to_float(line[24:32].rstrip()), string_to_float(line[32:40].rstrip())])
            if line.find("CTRIA3") != -1:
                elements.append([int(line[24:32].lstrip()) - 1, int(line[32:40].lstrip()) - 1,
                                 int(line[40:48].lstrip()) - 1])
    return nodes, elements


def _nastran_field(v):
    s = repr(float(v))
    if len(s) <= 8:
        return s.rjust(8)
    for prec in range(7, 0, -1):
        s = f"{v:.{prec}f}"
        if len(s) <= 8:
            return s.rjust(8)
    raise ValueError("coordinate does not fit a Nastran short field")


def write_fem(path, nodes, elements, scale=1.0):
    """Write a minimal OptiStruct deck the reference's parser (and parse_fem) reads back."""
    with open(path, "w", newline="") as f:
        f.write("$$ OptiStruct deck written by ceit_b200.readmesh.write_fem\nBEGIN BULK\n$$  GRID Data\n")
        for k, (x, y) in enumerate(np.asarray(nodes) * scale):
            f.write("GRID    %8d        %s%s     0.0\n" % (k + 1, _nastran_field(x), _nastran_field(y)))
        f.write("$$  CTRIA3 Data\n")
        for k, e in enumerate(np.asarray(elements)):
            f.write("CTRIA3  %8d%8d%8d%8d%8d\n" % (k + 1, 0, e[0] + 1, e[1] + 1, e[2] + 1))
        f.write("ENDDATA\n")


class ReadMesh(object):
    """DO NOT DIRECTLY USE THIS CLASS - call init_mesh() (readmesh.py:34-90)."""

    def __init__(self, filename, electrode_centers, electrode_radius, unit_name, folder_name="",
                 optimize_node_num=False, shuffle_element=False):
        self.config = get_config()
        self.folder_name = folder_name
        self.nodes, self.elements = parse_fem(
            self.config["rootdir"] + "/" + self.config["folder_name"] + "/" + filename)
        self.electrode_centers = electrode_centers
        self.electrode_num = len(self.electrode_centers)
        self.electrode_radius = electrode_radius
        unit_dict = {"m": 1, "cm": 0.1, "mm": 0.001, "inch": 0.0254}   # readmesh.py:67 (cm entry as is)
        transfer_param = 0
        if isinstance(unit_name, str):
            if unit_name not in unit_dict.keys():
                raise ValueError("Mesh unit name not inside the scope, please use one of mm, cm, m, inch.")
            transfer_param = unit_dict[unit_name]
        elif isinstance(unit_name, float):
            transfer_param = unit_name
        self.nodes = np.array(self.nodes) * transfer_param
        max_side = np.max(self.nodes) - np.min(self.nodes)
        if max_side < 0.05:
            print("The mesh seems to be too small, please make sure the length is accurate. "
                  "The max side length is: " + str(max_side) + "meter")
        self.clean_mesh()
        if optimize_node_num:
            self.optimize_node_number()
        if shuffle_element:
            self.shuffle_elements()
        self.output()
        self.out_2pkl()
        print('CSV and PKL file is ready. Mesh parsed.')

    def clean_mesh(self):
        """Merge nodes with identical coordinates (same result as readmesh.py:92-113)."""
        first = {}
        remap = np.zeros(len(self.nodes), dtype=np.int64)
        new_nodes = []
        for i, node in enumerate(self.nodes):
            key = (float(node[0]), float(node[1]))
            j = first.get(key)
            if j is None:
                j = len(new_nodes)
                first[key] = j
                new_nodes.append(node)
            remap[i] = j
        self.elements = [[int(remap[n]) for n in e] for e in self.elements]
        self.nodes = new_nodes

    def return_mesh(self):
        element_num = len(self.elements)
        mesh_obj = {'element': np.array(self.elements), 'node': np.array(self.nodes),
                    'perm': np.ones(element_num)}
        return mesh_obj, self.electrode_num, self.electrode_centers, self.electrode_radius

    def optimize_node_number(self):
        """Renumber nodes ring by ring from the outside in, by polar angle (readmesh.py:121-153)."""
        nodes = np.asarray(self.nodes)
        n = len(nodes)
        rad = np.sqrt(nodes[:, 0] ** 2 + nodes[:, 1] ** 2)
        with np.errstate(invalid="ignore", divide="ignore"):
            ang = np.where(nodes[:, 1] >= 0, np.arccos(nodes[:, 0] / rad), 2 * np.pi - np.arccos(nodes[:, 0] / rad))
        p_max = rad.max()
        steps = int(n / 10)
        p_inc = p_max / steps
        new_num = np.zeros(n, dtype=int)
        count = 0
        for i in range(steps + 1):
            sel = np.nonzero(((p_max - i * p_inc) >= rad) & (rad > (p_max - (i + 1) * p_inc)))[0]
            for k in sel[np.argsort(ang[sel], kind="stable")]:
                new_num[k] = count
                count += 1
        new_nodes = np.zeros((n, 2))
        new_nodes[new_num] = nodes
        self.elements = new_num[np.asarray(self.elements)]
        self.nodes = new_nodes

    def shuffle_elements(self):
        """Random element order (readmesh.py:155-164)."""
        N = len(self.elements)
        self.elements = [list(e) for e in self.elements]
        for i in range(N):
            r = np.random.randint(0, i + 1)
            self.elements[r], self.elements[i] = self.elements[i], self.elements[r]

    def output(self, file_name='Mesh_Cache'):
        base = self.config["rootdir"] + "/" + self.config["folder_name"] + "/" + file_name
        with open(base + '_Node.csv', mode='w', newline='') as outfile:
            w = csv.writer(outfile, delimiter=',')
            for node in self.nodes:
                w.writerow(node)
        with open(base + '_Element.csv', mode='w', newline='') as outfile_2:
            w = csv.writer(outfile_2, delimiter=',')
            for element in self.elements:
                w.writerow(element)

    def out_2pkl(self, filename='Mesh_'):
        d = self.config["rootdir"] + "/" + self.config["folder_name"]
        save_parameter(self.nodes, filename + 'nodes', d)
        save_parameter(self.elements, filename + 'elements', d)


class read_mesh_from_csv(object):
    """Read the mesh cache: mode 'csv' or 'pkl' (readmesh.py:191-242)."""

    def __init__(self, name='Mesh_Cache', mode='pkl'):
        self.config = get_config()
        self.folder_name = self.config["folder_name"]
        d = self.config["rootdir"] + "/" + self.config["folder_name"]
        if mode == 'csv':
            self.nodes, self.elements = [], []
            with open(d + '/' + name + '_Node.csv', newline='') as datafile:
                for line in csv.reader(datafile, delimiter=','):
                    if line:
                        self.nodes.append([float(x) for x in line])
            with open(d + '/' + name + '_Element.csv', newline='') as datafile_2:
                for line in csv.reader(datafile_2, delimiter=','):
                    if line:
                        self.elements.append([int(x) for x in line])
        elif mode == 'pkl':
            self.read_from_pkl()
        else:
            raise Exception('No such mesh reading mode.')
        self.electrode_centers = np.array(self.config["electrode_centers"])
        self.electrode_num = len(self.electrode_centers)
        self.electrode_radius = self.config["electrode_radius"]

    def read_from_pkl(self, filename='Mesh_'):
        d = self.config["rootdir"] + "/" + self.config["folder_name"]
        self.nodes = read_parameter(filename + 'nodes', d)
        self.elements = read_parameter(filename + 'elements', d)

    def return_mesh(self):
        element_num = len(self.elements)
        mesh_obj = {'element': np.array(self.elements), 'node': np.array(self.nodes),
                    'perm': np.ones(element_num)}
        return mesh_obj, self.electrode_num, self.electrode_centers, self.electrode_radius


def write_mesh_cache(folder, nodes, elements):
    """Write the pkl + csv cache for arrays that did not come from a .fem file (synthetic meshes)."""
    os.makedirs(folder, exist_ok=True)
    nodes_l = [np.array(n, dtype=np.float64) for n in np.asarray(nodes)]
    elems_l = [[int(a) for a in e] for e in np.asarray(elements)]
    save_parameter(nodes_l, 'Mesh_nodes', folder)
    save_parameter(elems_l, 'Mesh_elements', folder)
    with open(os.path.join(folder, 'Mesh_Cache_Node.csv'), 'w', newline='') as f:
        w = csv.writer(f, delimiter=',')
        for n in nodes_l:
            w.writerow(n)
    with open(os.path.join(folder, 'Mesh_Cache_Element.csv'), 'w', newline='') as f:
        w = csv.writer(f, delimiter=',')
        for e in elems_l:
            w.writerow(e)


def init_mesh(draw=False):
    """Run this after making a new mesh (readmesh.py:245-263). `draw` needs matplotlib."""
    config = get_config()
    read_mesh = ReadMesh(config["mesh_filename"], config["electrode_centers"], config["electrode_radius"],
                         config["mesh_unit"], config["folder_name"], config["optimize_node_num"],
                         config["shuffle_element"])
    mesh_obj, electrode_num, electrode_centers, electrode_radius = read_mesh.return_mesh()
    if draw:
        raise NotImplementedError("plotting is outside the scope of ceit_b200 (SURVEY.md §2 #8)")
    return mesh_obj, electrode_num, electrode_centers, electrode_radius
