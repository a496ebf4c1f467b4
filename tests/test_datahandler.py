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


def test_filename_helpers(dh):
    from ceit_b200.datahandler.DataHandler import get_corr_from_filename, get_height_idx_from_filename
    assert tuple(get_corr_from_filename("data_sample08_ex00100_20_35_50.csv")) == tuple(dh["name_corr"])
    assert get_height_idx_from_filename("calibration_sample01_ex00100_3.csv") == int(dh["name_height"])


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
    # a "recording" is anything with a delta_V attribute (the reference's Data after calc_delta_V)
    items = [types.SimpleNamespace(delta_V=dh["delta_V"][k]) for k in range(len(dh["raw"]))]
    with pytest.raises(Exception, match="WRONG!"):
        plotter.get_COP(types.SimpleNamespace(delta_V=None))
    for k, d in enumerate(items):
        x, y, v = plotter.get_COP(d)
        assert abs(x - dh["cop"][k, 0]) < 1e-12 and abs(y - dh["cop"][k, 1]) < 1e-12
        assert abs(v - dh["cop"][k, 2]) <= 1e-9 * abs(dh["cop"][k, 2])
    x, y, v, cnt = plotter.get_COP_batch(np.stack([d.delta_V for d in items], axis=1))
    assert np.allclose(x, dh["cop"][:, 0], rtol=0, atol=1e-12) and np.allclose(y, dh["cop"][:, 1], rtol=0, atol=1e-12)
    assert np.allclose(v, dh["cop"][:, 2], rtol=1e-9, atol=0) and np.all(cnt > 0)
    with pytest.raises(ZeroDivisionError):                    # delta_V == 0 -> constant (zero) map
        plotter.get_COP(types.SimpleNamespace(delta_V=np.zeros_like(dh["delta_V"][0])))
    import CEIT.datahandler.DataHandler as shim               # the reference's module path
    assert shim.DataPlotter is DataPlotter and shim.THRESHOLD_PERCENTAGE == 0.9
