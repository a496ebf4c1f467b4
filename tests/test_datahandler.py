"""datahandler row: recordings -> delta_V -> reconstruction -> centre of position.

CPU part: the oracle restatement and the host classes against vectors produced by the UNMODIFIED reference
(tests/golden/ref_datahandler.npz, generator: tests/golden/make_golden_datahandler.py).
GPU part: the batched reduction kernel (ceit_center_of_position) and the class pipeline on the device."""
import os
import sys
import types

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ceit_oracle as o  # noqa: E402

from ceit_b200 import meshgen  # noqa: E402
from ceit_b200.datahandler.Data import CalibData, Data, rev_translate_corr, translate_corr  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def dh():
    return dict(np.load(os.path.join(GOLDEN, "ref_datahandler.npz")))


def test_oracle_recording_mean_and_delta_v(dh):
    L = int(dh["L"])
    cm = o.recording_mean(dh["calib_raw"], L, True)
    assert cm.shape == (L * (L - 1),) and np.array_equal(cm, dh["calib_mean"])
    for k in range(len(dh["raw"])):
        assert np.array_equal(o.recording_mean(dh["raw"][k], L, True) - cm, dh["delta_V"][k])
    assert np.array_equal(o.recording_mean(dh["plain_data"]) - o.recording_mean(dh["plain_calib"]), dh["plain_delta_V"])


def test_oracle_center_of_position_matches_reference(golden, dh):
    m, r = golden("21_21")
    for k in range(len(dh["raw"])):
        cap = np.dot(r["inv_mat_289"], dh["delta_V"][k])
        got = o.center_of_position(cap, m["detection_index"], m["elem_param"], float(dh["threshold_percentage"]))
        assert np.array_equal(np.array(got), dh["cop"][k])
    with pytest.raises(ZeroDivisionError):
        o.center_of_position(np.ones(m["detection_index"].shape[0]), m["detection_index"], m["elem_param"])


def test_host_data_classes_match_reference(dh):
    calib = CalibData(dh["calib_raw"], 0, contain_excitation=True)
    assert np.array_equal(calib.data_mean, dh["calib_mean"]) and calib.height_idx == 0
    for k in range(len(dh["raw"])):
        d = Data(dh["raw"][k], int(dh["x"][k]), int(dh["y"][k]), int(dh["z"][k]), contain_excitation=True)
        assert (d.x, d.y, d.z) == (dh["tx"][k], dh["ty"][k], dh["z"][k])
        assert d.delta_V is None
        d.calc_delta_V(calib)
        assert np.array_equal(d.delta_V, dh["delta_V"][k])
        c = d.return_copy()
        assert (c.x, c.y, c.z) == (d.x, d.y, d.z) and np.array_equal(c.data_mean, d.data_mean) and c.delta_V is None
    d2, c2 = Data(dh["plain_data"], 10, 20, 30), CalibData(dh["plain_calib"], 1)
    d2.calc_delta_V(c2)
    assert np.array_equal(d2.delta_V, dh["plain_delta_V"])
    short = Data(dh["plain_data"][:, :100], 0, 0, 0)
    with pytest.raises(Exception, match="Shape does not match with calibration."):
        short.calc_delta_V(c2)
    assert translate_corr(20, 35) == (-80, 65) and rev_translate_corr(-80, 65) == (20, 35)
    # get_capacitance: empty without delta_V, cached after the first solve (Data.py:94-112)
    calls = []
    solver = types.SimpleNamespace(solve=lambda dv: calls.append(1) or np.arange(3.0))
    fresh = Data(dh["plain_data"], 1, 2, 3)
    assert fresh.get_capacitance(solver).size == 0 and not calls
    fresh.calc_delta_V(c2)
    assert np.array_equal(fresh.get_capacitance(solver), np.arange(3.0))
    assert np.array_equal(fresh.get_capacitance(solver), np.arange(3.0)) and len(calls) == 1


def test_filenames_and_dataset_layout(tmp_path, monkeypatch, dh):
    from ceit_b200.datahandler.DataHandler import DataHandler, get_corr_from_filename, get_height_idx_from_filename
    from ceit_b200.datahandler.Dataset import DataSet
    assert tuple(get_corr_from_filename("data_sample08_ex00100_20_35_50.csv")) == tuple(dh["name_corr"])
    assert get_height_idx_from_filename("calibration_sample01_ex00100_3.csv") == int(dh["name_height"])
    folder = tmp_path / "data" / "Smp08_001" / "0"
    folder.mkdir(parents=True)
    heights = [30, 50]
    np.savetxt(folder / "calibration_sample08_ex00100_0.csv", dh["calib_raw"], delimiter=",")
    np.savetxt(folder / "calibration_sample08_ex00100_1.csv", dh["calib_raw"] * 2, delimiter=",")
    np.savetxt(folder / "data_sample08_ex00100_20_35_50.csv", dh["raw"][0], delimiter=",")
    np.savetxt(folder / "data_sample08_ex00100_120_60_30.csv", dh["raw"][1], delimiter=",")
    monkeypatch.chdir(tmp_path)
    solver = types.SimpleNamespace()                       # a handler with a solver never builds one itself
    ds = DataSet("08", "001", 0, heights, solver=solver)
    assert sorted(ds.calibration_data) == heights and [len(ds.dataset[h]) for h in heights] == [1, 1]
    d = ds.get_data_at(-80, 65, 50)
    assert np.allclose(d.data, dh["raw"][0], rtol=0, atol=1e-15) and d.contain_excitation
    assert np.allclose(ds.get_calibration_data_at(50).data, 2 * dh["calib_raw"])
    assert ds.return_path_for_specific_point(-80, 65, 50) == (
        "./data/Smp08_001/0", "calibration_sample08_ex00100_1.csv", "data_sample08_ex00100_20_35_50.csv")
    with pytest.raises(IndexError):
        ds[40]
    with pytest.raises(IndexError):
        ds.get_data_at(0, 0, 50)
    np.savetxt(folder / "data_sample08_ex00100_1_1_99.csv", dh["raw"][2], delimiter=",")
    with pytest.raises(Exception, match="height list"):
        DataSet("08", "001", 0, heights, solver=solver)
    h = DataHandler(str(folder), solver=solver)
    assert h.read_one_calibration_file("calibration_sample08_ex00100_1.csv").height_idx == 1


# ---------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def torch_cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device; there is no CPU fallback")
    return torch


@pytest.mark.gpu
@pytest.mark.parametrize("n_det,F", [(722, 1), (722, 33), (4050, 257), (37, 64)])
def test_center_of_position_kernel_vs_oracle(torch_cuda, n_det, F):
    torch = torch_cuda
    from ceit_b200 import _lib
    rng = np.random.default_rng(n_det + F)
    maps = rng.standard_normal((n_det, F)) * np.exp(rng.random((n_det, 1)) * 4)
    if F > 2:
        maps[:, 2] = 3.25                                   # constant map: nothing above the threshold
    par = np.zeros((n_det, 9))
    par[:, 7:9] = rng.random((n_det, 2)) - 0.5
    idx = np.arange(n_det)
    got = _lib.center_of_position(torch.tensor(maps, device="cuda"), torch.tensor(par[:, 7].copy(), device="cuda"),
                                  torch.tensor(par[:, 8].copy(), device="cuda"), 0.9).cpu().numpy()
    for f in range(F):
        if F > 2 and f == 2:
            assert got[f, 3] == 0 and np.all(np.isnan(got[f, :3]))
            continue
        ref = o.center_of_position(maps[:, f], idx, par, 0.9)
        th = maps[:, f].max() * 0.9 + maps[:, f].min() * 0.1
        assert got[f, 3] == np.count_nonzero(maps[:, f] > th)                     # integer work: exact
        assert np.allclose(got[f, :3], ref, rtol=1e-13, atol=1e-15)               # sums in a different order


@pytest.mark.gpu
def test_cop_pipeline_on_device_vs_reference(torch_cuda, golden, dh, tmp_path, monkeypatch):
    m, r = golden("21_21")
    cfg = dict(meshgen.REFERENCE_CONFIG)
    cfg["folder_name"] = "MESH_21_21"
    path = meshgen.make_workspace(str(tmp_path / "w"), m["nodes"], m["elem"], cfg)
    monkeypatch.setenv("CEIT_B200_CONFIG", path)
    monkeypatch.chdir(os.path.dirname(path))
    np.save(os.path.join(os.path.dirname(path), "MESH_21_21", "JAC_cache.npy"), r["JAC"])
    from ceit_b200.Solver import Solver
    from ceit_b200.datahandler.DataHandler import DataPlotter
    plotter = DataPlotter(Solver(289))
    calib = CalibData(dh["calib_raw"], 0, contain_excitation=True)
    items = [Data(dh["raw"][k], int(dh["x"][k]), int(dh["y"][k]), int(dh["z"][k]), contain_excitation=True)
             for k in range(len(dh["raw"]))]
    with pytest.raises(Exception, match="WRONG!"):
        plotter.get_COP(items[0])
    for k, d in enumerate(items):
        d.calc_delta_V(calib)
        x, y, v = plotter.get_COP(d)
        assert abs(x - dh["cop"][k, 0]) < 1e-12 and abs(y - dh["cop"][k, 1]) < 1e-12
        assert abs(v - dh["cop"][k, 2]) <= 1e-9 * abs(dh["cop"][k, 2])
    tab = plotter.cop_table(items, calib)
    assert np.allclose(tab["x_re"], dh["cop"][:, 0] * 2000, rtol=0, atol=1e-9)
    assert np.allclose(tab["y_re"], dh["cop"][:, 1] * 2000, rtol=0, atol=1e-9)
    assert np.allclose(tab["values"], dh["cop"][:, 2], rtol=1e-9, atol=0)
    assert np.array_equal(tab["x_gt"], dh["tx"]) and np.array_equal(tab["y_gt"], dh["ty"])
    const = Data(np.tile(calib.data, (1, 1)), 0, 0, 0, contain_excitation=True)
    const.calc_delta_V(calib)                                # delta_V == 0 -> constant (zero) map
    with pytest.raises(ZeroDivisionError):
        plotter.get_COP(const)
    import CEIT.datahandler.DataHandler as shim               # the reference's module path
    assert shim.DataPlotter is DataPlotter and shim.THRESHOLD_PERCENTAGE == 0.9
