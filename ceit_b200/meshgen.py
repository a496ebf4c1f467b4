"""Synthetic structured sensor meshes and scratch workspaces.

The reference ships two HyperMesh decks only; its larger configurations are synthetic:
n x n nodes on a square sheet, pitch h (default 2 mm, the MESH_50_50 element size, so the
second-order signal J-1 stays ~1e-8 instead of vanishing as A^2), every cell split along the
(ix,iy)-(ix+1,iy+1) diagonal, nodes numbered row-major, square electrodes centred
`electrode_inset` inside the edge, `per_side` positions per side with shared corners
(5 -> 16 electrodes like config.json, 9 -> 32), listed counter-clockwise from (+a, 0).
"""
import json
import os

import numpy as np

from .readmesh import write_mesh_cache


def structured_mesh(n, pitch=2e-3):
    """nodes f64[n*n,2] (metres, centred), elem int64[2(n-1)^2,3] CCW."""
    c = (np.arange(n) - (n - 1) / 2.0) * pitch
    X, Y = np.meshgrid(c, c)                      # row-major: node = iy*n + ix
    nodes = np.stack([X.reshape(-1), Y.reshape(-1)], 1)
    ix, iy = np.meshgrid(np.arange(n - 1), np.arange(n - 1))
    n00 = (iy * n + ix).reshape(-1)
    n10, n01, n11 = n00 + 1, n00 + n, n00 + n + 1
    elem = np.empty((2 * len(n00), 3), dtype=np.int64)
    elem[0::2] = np.stack([n00, n10, n11], 1)
    elem[1::2] = np.stack([n00, n11, n01], 1)
    return nodes, elem


def perimeter_electrodes(half, per_side):
    """Centres on the square |x|,|y| = half, CCW from (half, 0); per_side positions per side."""
    t = np.linspace(-half, half, per_side)
    right = [(half, v) for v in t]                       # bottom -> top
    top = [(v, half) for v in t[::-1]][1:]               # right -> left (corner shared)
    left = [(-half, v) for v in t[::-1]][1:]
    bottom = [(v, -half) for v in t][1:-1]
    ring = right + top + left + bottom
    start = min(range(len(ring)), key=lambda k: (abs(ring[k][1]), -ring[k][0]))
    ring = ring[start:] + ring[:start]
    return [[float(x), float(y)] for x, y in ring]


def synthetic_config(n, per_side=5, pitch_mm=2.0, radius_mm=3.0, inset_mm=3.0, folder="MESH_SYN", **overrides):
    half_mm = (n - 1) * pitch_mm / 2.0
    centers = perimeter_electrodes(half_mm - inset_mm, per_side)
    cfg = {
        "mesh_filename": "EITMesh.fem", "optimize_node_num": False, "shuffle_element": False,
        "electrode_centers": centers, "electrode_radius": radius_mm, "signal_frequency": 20000,
        "resistance": 1, "reconstruction_mode": "n", "variable_change_for_JAC": 1e-3,
        "overall_origin_variable": 0, "detection_bound": 0.9 * half_mm, "is_first_JAC_calculation": True,
        "folder_name": folder, "calc_from": 0, "calc_end": len(centers), "regularization_coeff": 295,
        "device": "gpu", "sensor_param_unit": "mm", "mesh_unit": "mm",
    }
    cfg.update(overrides)
    return cfg


def make_workspace(root, nodes, elem, cfg):
    """Write <root>/config.json and the mesh cache under <root>/<folder_name>; returns config path."""
    os.makedirs(root, exist_ok=True)
    write_mesh_cache(os.path.join(root, cfg["folder_name"]), nodes, elem)
    path = os.path.join(root, "config.json")
    with open(path, "w") as f:
        json.dump(cfg, f, indent=1)
    return path


def make_synthetic_workspace(root, n, per_side=5, **kw):
    cfg = synthetic_config(n, per_side, **kw)
    nodes, elem = structured_mesh(n, kw.get("pitch_mm", 2.0) * 1e-3)
    return make_workspace(root, nodes, elem, cfg)


def workspace_from_arrays(root, nodes, elem, reference_like_cfg):
    """Workspace for a shipped mesh held as arrays (tests/golden/mesh_*.npz)."""
    return make_workspace(root, nodes, elem, reference_like_cfg)


REFERENCE_CONFIG = {
    "mesh_filename": "EITMesh.fem", "optimize_node_num": False, "shuffle_element": False,
    "electrode_centers": [[47, 0], [47, 23.5], [47, 47], [23.5, 47], [0, 47], [-23.5, 47], [-47, 47],
                          [-47, 23.5], [-47, 0], [-47, -23.5], [-47, -47], [-23.5, -47], [0, -47],
                          [23.5, -47], [47, -47], [47, -23.5]],
    "electrode_radius": 3, "signal_frequency": 20000, "resistance": 1, "reconstruction_mode": "n",
    "variable_change_for_JAC": 1e-3, "overall_origin_variable": 0, "detection_bound": 45,
    "is_first_JAC_calculation": True, "folder_name": "MESH_50_50", "calc_from": 0, "calc_end": 16,
    "regularization_coeff": 295, "device": "gpu", "sensor_param_unit": "mm", "mesh_unit": "mm",
}
