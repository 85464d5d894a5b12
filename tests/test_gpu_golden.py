"""CUDA path against the committed golden vectors (tests/golden/*.npz) -- no oracle run needed on the box."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    "hd16_rk4_f32": ("hd", "RK4", np.float32, 5, 5e-3),
    "mhd16_rk4_f32": ("mhd", "RK4", np.float32, 5, 5e-3),
    "mhd16_lsrk54_f64": ("mhd", "LSRK54", np.float64, 5, 5e-3),
    "emhd16_rk4_f64": ("emhd", "RK4", np.float64, 5, 2e-4),
    "mhd16_tg_rk4_f32": ("mhd", "RK4", np.float32, 10, 1e-2),
}


def rel(a, b):
    return float(np.linalg.norm((a - b).ravel()) / np.linalg.norm(b.ravel()))


@pytest.mark.parametrize("name", sorted(CASES))
def test_cuda_matches_golden(name):
    import mhdflows_jl_b200 as M
    kind, stepper, T, nsteps, dt = CASES[name]
    ref = np.load(os.path.join(HERE, "golden", name + ".npz"))
    kw = dict(nx=16, T=T, dt=dt, stepper=stepper)
    if kind == "mhd":
        kw.update(nu=2e-2, eta=3e-2, B_field=True)
    elif kind == "hd":
        kw.update(nu=2e-2)
    else:
        kw.update(B_field=True, EMHD=True)
    p = M.Problem(M.GPU(), **kw)
    M.SetUpProblemIC(p, **{k[3:]: ref[k] for k in ref.files if k.startswith("ic_")})
    msk = p.grid.retained_mask()
    tol = 1e-5 if T is np.float32 else 1e-12
    N0 = p.calcN()
    for i in range(p.Nl):
        if np.linalg.norm(ref["N0"][i]) > 1e-3 * np.linalg.norm(ref["N0"]):
            assert rel(N0[i][msk], ref["N0"][i]) < tol, (name, "N0", i)
    # calcN refreshed the stale vars exactly like the reference's calcN! does; state is untouched
    energies = []
    for _ in range(nsteps):
        M.stepforward(p)
        ke, me = p.energy(M.STALE)
        energies.append((me,) if kind == "emhd" else ((ke, me) if kind == "mhd" else (ke,)))
    sol = p.sol
    for i in range(p.Nl):
        assert rel(sol[i][msk], ref["sol"][i]) < tol, (name, "sol", i)
    assert np.allclose(np.array(energies), ref["energies"], rtol=20 * tol)
    assert abs(p.clock.t - float(ref["t"])) < 1e-6
    p.close()
