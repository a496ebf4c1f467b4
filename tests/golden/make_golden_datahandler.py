#!/usr/bin/env python
"""Golden vectors for the datahandler row of SURVEY.md §8(f): RUN THE UNMODIFIED REFERENCE classes
`CEIT.datahandler.Data.{Data, CalibData}` and `CEIT.datahandler.DataHandler.DataPlotter.get_COP`
(DataHandler.py:28-56) on seeded synthetic recordings and store inputs + outputs.

Build container only (needs /root/reference).  The solver handed to DataPlotter is a stand-in that does what
Solver.solve does (np.dot(inv_mat, delta_V), Solver.py:95-108) with the golden MESH_21_21 inverse model, so the
vectors pin Data / CalibData (row mean, excitation columns removed), calc_delta_V and the thresholded centre of
position, not the inverse model (pinned elsewhere).

    python tests/golden/make_golden_datahandler.py [--ref /root/reference]
Output: tests/golden/ref_datahandler.npz
"""
import argparse
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    args = ap.parse_args()
    mesh = np.load(os.path.join(HERE, "mesh_21_21.npz"))
    ref = np.load(os.path.join(HERE, "ref_21_21.npz"))
    cwd = os.getcwd()
    tree = mg._make_tree(args.ref, "MESH_21_21")
    mg._install_shims()
    sys.modules["matplotlib.pyplot"].axes = object          # only used as a type annotation (DataHandler.py:96)
    mg._import_ref(tree)
    from CEIT.datahandler.Data import CalibData, Data
    from CEIT.datahandler import DataHandler as DH

    solver = types.SimpleNamespace()
    inv_mat = ref["inv_mat_289"]
    solver.solve = lambda dV: np.dot(inv_mat, dV)
    solver.mesh = types.SimpleNamespace(detection_index=mesh["detection_index"], elem_param=mesh["elem_param"])
    plotter = DH.DataPlotter(solver=solver)

    rng = np.random.default_rng(20240607)
    L, S, n_items = 16, 12, 24
    m = L * (L - 1)
    # recordings WITH excitation channels (L*L columns, contain_excitation=True) and without (m columns)
    calib_raw = 1.0 + 1e-3 * rng.standard_normal((S, L * L))
    calib = CalibData(calib_raw, 0, contain_excitation=True)
    out = {"calib_raw": calib_raw, "calib_mean": calib.data_mean, "L": L}
    raws, xs, ys, zs, dVs, cops, tx, ty = [], [], [], [], [], [], [], []
    for k in range(n_items):
        # a localised bump in measurement space: baseline + response of a random detection element
        col = rng.integers(0, inv_mat.shape[0])
        resp = np.zeros(L * L)
        keep = [i for i in range(L * L) if i % L != 0]
        resp[keep] = 1e-2 * (ref["JAC"][:, mesh["detection_index"][col]] - 1.0) / 1e-8
        if k >= n_items // 2:                                      # arbitrary responses: several elements pass the threshold
            resp[keep] = 1e-3 * rng.standard_normal(m) * np.exp(-rng.random(m) * 3)
        raw = calib_raw.mean(axis=0) + resp + 1e-5 * rng.standard_normal((S, L * L))
        x, y, z = int(rng.integers(0, 200)), int(rng.integers(0, 200)), int(rng.choice([30, 50]))
        d = Data(raw, x, y, z, contain_excitation=True)
        d.calc_delta_V(calib)
        cops.append(plotter.get_COP(d))
        raws.append(raw); xs.append(x); ys.append(y); zs.append(z); dVs.append(d.delta_V); tx.append(d.x); ty.append(d.y)
    d2 = Data(rng.random((S, m)), 10, 20, 30)                       # no excitation channels: plain mean
    c2 = CalibData(rng.random((S, m)), 1)
    d2.calc_delta_V(c2)
    out.update(raw=np.array(raws), x=np.array(xs), y=np.array(ys), z=np.array(zs), delta_V=np.array(dVs),
               cop=np.array(cops), tx=np.array(tx), ty=np.array(ty), plain_data=d2.data, plain_calib=c2.data,
               plain_delta_V=d2.delta_V, threshold_percentage=DH.THRESHOLD_PERCENTAGE,
               name_corr=np.array(DH.get_corr_from_filename("data_sample08_ex00100_20_35_50.csv")),
               name_height=DH.get_height_idx_from_filename("calibration_sample01_ex00100_3.csv"))
    os.chdir(cwd)
    np.savez_compressed(os.path.join(HERE, "ref_datahandler.npz"), **out)
    print("wrote ref_datahandler.npz:", {k: np.shape(v) for k, v in out.items()})
    print("cop[0..2]", np.array(cops)[:3])


if __name__ == "__main__":
    main()
