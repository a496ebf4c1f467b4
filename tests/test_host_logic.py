"""Host-side logic and the C ABI surface, CPU only (no compute calls)."""
import ctypes
import json
import os
import re

import numpy as np
import pytest

from ceit_b200 import _lib, meshgen
from ceit_b200.sharding import all_shards

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture()
def workspace(tmp_path, golden, monkeypatch):
    def make(tag, **overrides):
        m, _ = golden(tag)
        cfg = dict(meshgen.REFERENCE_CONFIG)
        cfg["folder_name"] = "MESH_" + tag
        cfg.update(overrides)
        path = meshgen.make_workspace(str(tmp_path / tag), m["nodes"], m["elem"], cfg)
        monkeypatch.setenv("CEIT_B200_CONFIG", path)
        return path
    return make


def test_header_symbols_are_exported():
    hdr = open(os.path.join(ROOT, "include", "ceit_b200.h")).read()
    names = set(re.findall(r"\b(ceit_[a-z_0-9]+)\s*\(", hdr))
    assert {"ceit_plan_create", "ceit_jacobian_block", "ceit_inverse_model", "ceit_reconstruct"} <= names
    L = _lib.lib()
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/ceit_b200.h but not exported"
    assert set(_lib.SYMBOLS) <= names
    assert L.ceit_version() >= 100


def test_compute_entry_fails_loudly_without_plan_state():
    L = _lib.lib()
    assert L.ceit_factor(ctypes.c_void_p(0), 0, ctypes.c_void_p(0)) < 0
    assert b"null" in L.ceit_last_error()


@pytest.mark.parametrize("tag", ["21_21", "50_50"])
@pytest.mark.parametrize("target,nd", [(0, 40), (1, 40), (64, 0), (10 ** 6, 0), (0, 20), (192, 20)])
def test_plan_tables(golden, tag, target, nd):
    m, r = golden(tag)
    _lib.set_option("nd_interior", nd)
    try:
        p = _lib.Plan(m["nodes"], m["elem"], m["electrodes"], target_block=target, upload=False)
    finally:
        _lib.set_option("nd_interior", 64)
    tb = p.tables()
    # CSR pattern bit-exact with scipy.sparse.csr_matrix(reference K_sparse)
    assert np.array_equal(tb["rowptr"], r["K_indptr"]) and np.array_equal(tb["colidx"], r["K_indices"])
    N = p.N
    assert sorted(tb["pos"].tolist()) == list(range(N))                            # compact: a permutation
    assert np.array_equal(tb["pos"][tb["U"]], p.N0 + np.arange(p.nU))
    pos = tb["ppos"]
    assert len(set(pos.tolist())) == N and pos.min() >= 0 and pos.max() < p.NP     # padded: injective
    U = tb["U"]
    expect_U = []
    for e in m["electrodes"]:
        for n in np.unique(m["elem"][np.asarray(e)].reshape(-1)):
            if n not in expect_U:
                expect_U.append(int(n))
    assert U.tolist() == expect_U
    assert np.array_equal(pos[U], p.N0P + np.arange(p.nU))
    assert p.NP == p.N0P + p.nU
    if target == 0 and nd > 0:
        # default: multi-level nested dissection (the numerics of this plan are checked against a dense inverse in
        # tests/test_nd_tree_host.py); here only the shape of the tables
        assert p.tree_levels >= 3 and p.P == 0 and p.n_blocks == 0 and p.N0P >= p.N0
        p.close()
        return
    assert p.tree_levels == 0
    NI = p.P * p.nI
    assert p.N0P == NI + p.nC
    rows = np.repeat(np.arange(N), np.diff(tb["rowptr"]))
    pr, pc = pos[rows], pos[tb["colidx"]]
    # level 2: every edge between separator nodes joins the same or adjacent blocks
    sep = (pr >= NI) & (pr < p.N0P) & (pc >= NI) & (pc < p.N0P)
    blk = np.searchsorted(tb["blk_ptr"], np.arange(p.nC), side="right") - 1
    if sep.any():
        assert np.abs(blk[pr[sep] - NI] - blk[pc[sep] - NI]).max() <= 1
    # level 1: an interior node only touches its own interior, separators or electrode nodes
    inter = (pr < NI) & (pc < NI)
    if p.P:
        assert np.all(pr[inter] // p.nI == pc[inter] // p.nI)
        assert p.nI <= 64
    else:
        assert not inter.any()
    if nd == 0 and target == 10 ** 6:
        assert p.n_blocks == 1 and p.P == 0
    if nd > 0 and tag == "50_50":
        assert p.P > 4
    p.close()


def test_plan_rejects_bad_input(golden):
    m, _ = golden("21_21")
    bad = m["elem"].copy()
    bad[0, 0] = len(m["nodes"]) + 5
    with pytest.raises(_lib.CeitError):
        _lib.Plan(m["nodes"], bad, m["electrodes"], upload=False)
    with pytest.raises(_lib.CeitError, match="No element is selected"):
        _lib.Plan(m["nodes"], m["elem"], [m["electrodes"][0], []], upload=False)
    with pytest.raises(Exception, match="2D Node coordinates incorrect"):
        _lib.Plan(np.zeros((4, 3)), m["elem"], m["electrodes"], upload=False)


@pytest.mark.parametrize("tag", ["21_21", "50_50"])
def test_meshobj_matches_reference(workspace, golden, tag):
    workspace(tag)
    from ceit_b200.models.mesh import MeshObj
    m, _ = golden(tag)
    mesh = MeshObj()
    assert np.array_equal(mesh.elem, m["elem"]) and np.array_equal(mesh.nodes, m["nodes"])
    scale = np.abs(m["elem_param"]).max(axis=0)
    assert (np.abs(mesh.elem_param - m["elem_param"]) / scale).max() < 1e-12
    for i, e in enumerate(m["electrodes"]):
        assert mesh.electrode_mesh[i] == [int(x) for x in e]
    assert np.array_equal(mesh.detection_index, m["detection_index"])
    assert np.array_equal(mesh.detection_elem, m["detection_elem"])
    hull = mesh.get_perimeter()
    assert len(hull) >= 4 and len(set(hull.tolist())) == len(hull)


def test_get_config_semantics(workspace):
    path = workspace("21_21", sensor_param_unit="mm")
    from ceit_b200.util.utilities import get_config
    cfg = get_config()
    assert cfg["rootdir"] == os.path.dirname(path)
    assert abs(cfg["electrode_radius"] - 0.003) < 1e-18 and abs(cfg["detection_bound"] - 0.045) < 1e-18
    assert abs(cfg["electrode_centers"][1][1] - 0.0235) < 1e-18
    raw = json.load(open(path))
    raw["sensor_param_unit"] = "SI"
    json.dump(raw, open(path, "w"))
    assert get_config()["electrode_radius"] == 3          # re-read on every call, no conversion for SI


def test_fem_roundtrip_and_cache_formats(tmp_path, golden, monkeypatch):
    """write_fem -> init_mesh (parse, unit scale, duplicate merge, csv + pkl cache) -> read back."""
    m, _ = golden("21_21")
    from ceit_b200 import readmesh
    root = tmp_path / "ws"
    folder = root / "MESH_X"
    folder.mkdir(parents=True)
    nodes_mm = m["nodes"] * 1000.0
    # append a duplicate of node 3 and let one element use it: it must be merged away
    nodes_dup = np.concatenate([nodes_mm, nodes_mm[3:4]], 0)
    elem = m["elem"].copy()
    hit = np.argwhere(elem == 3)[0]
    elem[hit[0], hit[1]] = len(nodes_mm)
    readmesh.write_fem(str(folder / "EITMesh.fem"), nodes_dup, elem)
    cfg = dict(meshgen.REFERENCE_CONFIG)
    cfg["folder_name"] = "MESH_X"
    json.dump(cfg, open(root / "config.json", "w"))
    monkeypatch.setenv("CEIT_B200_CONFIG", str(root / "config.json"))
    mesh_obj, L, centers, radius = readmesh.init_mesh()
    assert L == 16 and mesh_obj["node"].shape == m["nodes"].shape
    assert np.array_equal(mesh_obj["element"], m["elem"])
    assert np.abs(mesh_obj["node"] - m["nodes"]).max() < 5e-7          # 8-char Nastran fields
    for mode in ("pkl", "csv"):
        mo, _, _, _ = readmesh.read_mesh_from_csv(mode=mode).return_mesh()
        assert np.array_equal(mo["element"], m["elem"])
        assert np.allclose(mo["node"], mesh_obj["node"], rtol=0, atol=1e-15)
    import pickle
    nodes_pkl = pickle.load(open(folder / "cache_Mesh_nodes.pkl", "rb"))
    elems_pkl = pickle.load(open(folder / "cache_Mesh_elements.pkl", "rb"))
    assert isinstance(nodes_pkl, list) and nodes_pkl[0].shape == (2,) and isinstance(elems_pkl[0], list)
    assert readmesh.string_to_float("1.5-3") == 1.5e-3 and readmesh.string_to_float("-2.5-2") == -2.5e-2


def test_reference_fem_parses_identically(golden):
    """When the reference tree is present (build container only), its decks parse to the goldens."""
    path = "/root/reference/MESH_21_21/EITMesh.fem"
    if not os.path.exists(path):
        pytest.skip("reference tree not present")
    from ceit_b200.readmesh import parse_fem
    m, _ = golden("21_21")
    n, e = parse_fem(path)
    assert np.array_equal(np.array(e), m["elem"]) and np.abs(np.array(n) * 0.001 - m["nodes"]).max() == 0


def test_synthetic_mesh_generator():
    nodes, elem = meshgen.structured_mesh(9)
    assert nodes.shape == (81, 2) and elem.shape == (128, 3)
    x, y = nodes[elem, 0], nodes[elem, 1]
    det = (x[:, 1] - x[:, 0]) * (y[:, 2] - y[:, 0]) - (x[:, 2] - x[:, 0]) * (y[:, 1] - y[:, 0])
    assert np.all(det > 0)                                             # all CCW
    assert meshgen.perimeter_electrodes(47, 5) == [[float(a), float(b)] for a, b in
                                                   meshgen.REFERENCE_CONFIG["electrode_centers"]]
    assert len(meshgen.perimeter_electrodes(100, 9)) == 32


@pytest.mark.parametrize("L,E,ws", [(16, 5000, 1), (16, 5000, 8), (16, 101, 3), (4, 10, 6), (32, 77, 8), (3, 50, 7)])
def test_shards_cover_every_column_once(L, E, ws):
    cover = np.zeros((L, E), dtype=int)
    for (i0, i1, e0, e1) in all_shards(L, E, ws):
        cover[i0:i1, e0:e1] += 1
    assert np.all(cover == 1)


def test_reference_module_paths_import():
    """CEIT.* re-exports (the reference's Example scripts import these paths)."""
    import importlib
    for mod, names in (("CEIT.EFEM", ["EFEM"]), ("CEIT.EJAC", ["EJAC"]), ("CEIT.Solver", ["Solver", "reinitialize_solver"]),
                       ("CEIT.readmesh", ["init_mesh", "read_mesh_from_csv"]), ("CEIT.models.mesh", ["MeshObj"]),
                       ("CEIT.util.utilities", ["get_config", "read_parameter", "save_parameter"]),
                       ("CEIT.Simulator", ["Simulator"]), ("CEIT.FEMBasic", ["FEMBasic"])):
        m = importlib.import_module(mod)
        for n in names:
            assert hasattr(m, n), (mod, n)


def test_cpu_device_is_rejected_and_no_gpu_fails_loudly(workspace):
    workspace("21_21", device="cpu")
    from ceit_b200.EFEM import EFEM
    with pytest.raises(Exception, match="no CPU path"):
        EFEM()
    workspace("21_21", device="tpu")
    with pytest.raises(Exception, match='specified device inside the config file'):
        EFEM()
    import torch
    if not torch.cuda.is_available():
        workspace("21_21")
        with pytest.raises(_lib.CeitError, match="no CPU fallback"):
            EFEM()


def test_bench_reference_arm_contract(capsys):
    """`bench.py --impl reference` prints one JSON line with the contract's keys (1 tiny step, CPU only)."""
    import argparse
    import bench
    bench.run_reference(argparse.Namespace(gpus=1, steps=1, warmup=0, impl="reference", config="50_50"))
    line = [l for l in capsys.readouterr().out.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["impl"] == "reference" and d["metric"] == "jacobian_columns_per_s" and d["unit"] == "columns/s"
    assert d["value"] > 0 and d["higher_is_better"] is True
    # "reference" when the git-ignored copy oracle/_ref exists (built where /root/reference is present), else the port
    ref_copy = os.path.isfile(os.path.join(ROOT, "oracle", "_ref", "CEIT", "EFEM.py"))
    assert d["cpu_baseline"]["kind"] == ("reference" if ref_copy else "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["config"]["workload"] == bench.WORKLOAD


def _quarter_conflicts(local, eptr):
    """wavefronts beyond the minimum of the Woodbury kernel's 16-byte row loads: per (quarter warp, corner) the number
    of DISTINCT slots that share a residue mod 8 with another distinct slot (woodbury_lr_pitch: pitch/2 odd)."""
    extra = groups = 0
    for p in range(len(eptr) - 1):
        loc = local[3 * eptr[p]:3 * eptr[p + 1]].reshape(-1, 3)
        for q in range(0, len(loc), 8):
            for a in range(3):
                slots = np.unique(loc[q:q + 8, a])
                extra += len(slots) - len(np.unique(slots % 8))
                groups += 1
    return extra, groups


@pytest.mark.parametrize("tag", ["21_21", "50_50"])
def test_woodbury_patches_cover_elements_and_slots_are_spread(golden, tag):
    """ceit_plan_patches: every element in exactly one patch, <= 32 elements and <= node_cap nodes per patch, the slot
    table consistent with the element connectivity, and the slot assignment of plan.cpp (build_patches) leaves few
    shared-memory bank conflicts between the rows a quarter warp reads together."""
    m, _ = golden(tag)
    p = _lib.Plan(m["nodes"], m["elem"], m["electrodes"], upload=False)
    tb, pt = p.tables(), p.patches()
    E = len(m["elem"])
    assert sorted(pt["elems"].tolist()) == list(range(E))
    ne = np.diff(pt["eptr"])
    nn = np.diff(pt["nptr"])
    assert ne.min() >= 1 and ne.max() <= 32 and nn.max() <= pt["node_cap"] <= 64
    pos = tb["pos"]
    for k in range(len(ne)):
        nodes = pt["nodes"][pt["nptr"][k]:pt["nptr"][k + 1]]
        assert len(set(nodes.tolist())) == len(nodes)                       # one slot per node
        loc = pt["local"][3 * pt["eptr"][k]:3 * pt["eptr"][k + 1]].reshape(-1, 3)
        assert loc.min() >= 0 and loc.max() < len(nodes)
        assert set(loc.reshape(-1).tolist()) == set(range(len(nodes)))      # every staged row is used
        el = pt["elems"][pt["eptr"][k]:pt["eptr"][k + 1]]
        assert np.array_equal(nodes[loc], pos[m["elem"][el]])               # slot -> node position of each corner
    extra, groups = _quarter_conflicts(pt["local"], pt["eptr"])
    # first-occurrence numbering of the same patches, for comparison
    first = pt["local"].copy()
    for k in range(len(ne)):
        seg = first[3 * pt["eptr"][k]:3 * pt["eptr"][k + 1]]
        order = {}
        for v in seg:
            order.setdefault(int(v), len(order))
        first[3 * pt["eptr"][k]:3 * pt["eptr"][k + 1]] = [order[int(v)] for v in seg]
    extra0, _ = _quarter_conflicts(first, pt["eptr"])
    assert extra <= extra0 and extra <= 0.25 * groups, (extra, extra0, groups)
    p.close()


def test_write_behind_npy_helpers(tmp_path, monkeypatch):
    """save_npy_async / wait_npy / load_npy (util/utilities.py): synchronous unless CEIT_B200_WRITE_BEHIND=1; with it
    the write runs on a worker thread, readers and rewriters of the same path join first, bytes == np.save"""
    from ceit_b200.util import utilities as u
    a = np.random.default_rng(0).random((240, 882))
    ref = tmp_path / "ref.npy"
    np.save(ref, a)
    path = str(tmp_path / "JAC_cache.npy")
    monkeypatch.delenv("CEIT_B200_WRITE_BEHIND", raising=False)
    u.save_npy_async(path, a)
    assert not u._pending and open(path, "rb").read() == open(ref, "rb").read()
    monkeypatch.setenv("CEIT_B200_WRITE_BEHIND", "1")
    for k in range(5):                                   # same size: overwritten in place, each time joined first
        b = a + k
        u.save_npy_async(path, b)
        assert np.array_equal(u.load_npy(path), b)       # load_npy joins the pending write
        assert not u._pending
    u.save_npy_async(path, a)
    u.save_npy_async(str(tmp_path / "inv_mat.npy"), a.T.copy())
    u.wait_npy()
    assert open(path, "rb").read() == open(ref, "rb").read()
    assert np.array_equal(np.load(tmp_path / "inv_mat.npy"), a.T)
    u.save_npy_async(str(tmp_path / "no_such_dir" / "x.npy"), a)
    with pytest.raises(OSError):
        u.wait_npy()                                     # the worker's error surfaces at the join
    assert not u._pending


@pytest.mark.parametrize("tag,parts", [("21_21", 2), ("50_50", 2), ("50_50", 4), ("50_50", 8)])
def test_domain_decomposed_plans_partition_the_elements(golden, tag, parts):
    """Multi-GPU plans (ceit_factor_phase): the ranks' element sets are a partition, every rank holds a strict subset of
    the node rows, the exchange buffers have the same shape on every rank, and every patch of a rank only touches rows
    that rank holds (host tables only, no device)."""
    m, _ = golden(tag)
    N, E = len(m["nodes"]), len(m["elem"])
    plans = [_lib.Plan(m["nodes"], m["elem"], m["electrodes"], upload=False, parts=parts, part=r) for r in range(parts)]
    own = np.stack([p.own_elements() for p in plans])
    assert (own.sum(axis=0) == 1).all()
    assert own.sum(axis=1).min() > 0
    infos = [p.exchange_info()[0] for p in plans]
    for r, (p, info) in enumerate(zip(plans, infos)):
        assert info[0] == parts and info[1] == r
        assert (info[2:4] == infos[0][2:4]).all() and info[6] == infos[0][6]      # same exchange shapes, same cut depth
        assert info[5] < p.N0 and info[4] == info[5] + p.nU                       # a strict subset of the F0 rows
        pos = p.tables()["pos"]
        held = pos >= 0
        assert held.sum() == info[4]
        assert held[np.asarray(m["elem"])[own[r]]].all()                          # nodes of the own elements are held
        pt = p.patches()
        n_el = int(pt["eptr"][-1])
        assert sorted(pt["elems"][:n_el].tolist()) == np.flatnonzero(own[r]).tolist()   # patches = exactly the own elements
        assert pt["nodes"].min() >= 0 and pt["nodes"].max() < info[4]
    # the union of the held non-electrode rows covers the mesh
    covered = np.zeros(N, bool)
    for p in plans:
        covered |= p.tables()["pos"] >= 0
    assert covered.all()
    with pytest.raises(_lib.CeitError):
        _lib.Plan(m["nodes"], m["elem"], m["electrodes"], upload=False, parts=2, part=2)
