"""Second-order scaling property of the finite-difference 'Jacobian' (SURVEY.md §7.1): on a real
base the perturbation is first-order imaginary, so J-1 scales like c^2.  GPU only."""
import numpy as np
import pytest

from ceit_b200 import _lib

pytestmark = pytest.mark.gpu


def test_signal_scales_quadratically(golden):
    import torch as t
    if not t.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device")
    m, r = golden("50_50")
    E, L = len(m["elem"]), 16
    omega = 2 * np.pi * 2e4
    p = _lib.Plan(m["nodes"], m["elem"], m["electrodes"])
    p.assemble(t.tensor(m["elem_perm"], device="cuda"), t.zeros(E, dtype=t.float64, device="cuda"), omega)
    p.factor(False)
    J1 = t.zeros((L * (L - 1), E), dtype=t.float64, device="cuda")
    J2 = t.zeros_like(J1)
    p.jacobian_block(0, L, 0, E, 1e-3, omega, J1, E)
    p.jacobian_block(0, L, 0, E, 2e-3, omega, J2, E)
    d1, d2 = (J1 - 1), (J2 - 1)
    big = d1.abs() > 1e-10
    ratio = (d2[big] / d1[big])
    assert (ratio - 4).abs().max().item() < 1e-2
    p.close()
