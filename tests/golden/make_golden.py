"""Generate the golden vectors under tests/golden/ from the CPU oracle (oracle/mhdflows_oracle.py).

    python tests/golden/make_golden.py

The reference (MHDFlows.jl) ships no tests or fixtures and Julia is not installable here, so these vectors pin
the ORACLE (literal restatement), not the Julia package itself ("parity unpinned").  They are
small (16^3, compressed, retained modes only) and let both the CPU suite (oracle regression) and the GPU suite
(CUDA path vs frozen numbers) run without regenerating anything.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import mhdflows_oracle as O  # noqa: E402

CASES = {
    # name: (kind, stepper, dtype, nsteps, dt, turbulent IC)
    "hd16_rk4_f32": ("hd", "RK4", np.float32, 5, 5e-3, True),
    "mhd16_rk4_f32": ("mhd", "RK4", np.float32, 5, 5e-3, True),
    "mhd16_lsrk54_f64": ("mhd", "LSRK54", np.float64, 5, 5e-3, True),
    "emhd16_rk4_f64": ("emhd", "RK4", np.float64, 5, 2e-4, True),
    "mhd16_tg_rk4_f32": ("mhd", "RK4", np.float32, 10, 1e-2, False),
}


def build(kind, stepper, T, dt, turb, n=16):
    kw = dict(nx=n, T=T, dt=dt, stepper=stepper)
    if kind == "mhd":
        p = O.Problem(nu=2e-2, eta=3e-2, B_field=True, **kw)
    elif kind == "hd":
        p = O.Problem(nu=2e-2, **kw)
    else:
        p = O.Problem(B_field=True, EMHD=True, **kw)
    g = p.grid
    if turb:
        u, b = O.random_phase_ic(g, 1234), O.random_phase_ic(g, 5678)
    else:
        ic = O.taylor_green_ic(g)
        u, b = ic[:3], ic[3:]
    if kind == "emhd":
        O.SetUpProblemIC(p, bx=b[0], by=b[1], bz=b[2])
        ic_fields = dict(bx=b[0], by=b[1], bz=b[2])
    elif kind == "mhd":
        O.SetUpProblemIC(p, *u, bx=b[0], by=b[1], bz=b[2])
        ic_fields = dict(ux=u[0], uy=u[1], uz=u[2], bx=b[0], by=b[1], bz=b[2])
    else:
        O.SetUpProblemIC(p, *u)
        ic_fields = dict(ux=u[0], uy=u[1], uz=u[2])
    return p, ic_fields


def run_case(name):
    kind, stepper, T, nsteps, dt, turb = CASES[name]
    p, ic = build(kind, stepper, T, dt, turb)
    g = p.grid
    msk = g.retained_mask()
    N0 = np.zeros_like(p.sol)
    q, _ = build(kind, stepper, T, dt, turb)
    q.calcN(N0, q.sol, 0.0, q.clock, q.vars, q.params, q.grid)
    energies = []
    for _ in range(nsteps):
        O.stepforward(p)
        e = O.ProbDiagnostic(p, rounded=False)
        energies.append(e if isinstance(e, tuple) else (e,))
    out = {"N0": np.stack([f[msk] for f in N0]), "sol": np.stack([f[msk] for f in p.sol]),
           "energies": np.array(energies, dtype=np.float64), "t": np.float64(p.clock.t)}
    out.update({"ic_" + k: v for k, v in ic.items()})
    return out


if __name__ == "__main__":
    for name in CASES:
        d = run_case(name)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
        print(name, {k: (v.shape, str(v.dtype)) for k, v in d.items() if hasattr(v, "shape")})
