"""Opt-in spectral kernel (MHDF_SPEC2=1) vs the default one: the spectral state after a few steps must be bit-identical
(same arithmetic in the same order) for HD / MHD / EMHD, RK4 and LSRK54.  Run under gpurun; prints one line per case."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mhdflows_jl_b200 as M  # noqa: E402


def run(flag, phys, stepper, n=32, steps=3):
    os.environ["MHDF_SPEC2"] = flag
    kw = dict(nx=n, ny=n, nz=n, T=np.float32, stepper=stepper)
    if phys == "hd":
        kw.update(nu=2e-2, dt=4e-3)
    elif phys == "mhd":
        kw.update(nu=2e-2, eta=3e-2, B_field=True, dt=4e-3)
    else:
        kw.update(B_field=True, EMHD=True, dt=2e-4)
    p = M.Problem(M.GPU(), **kw)
    rng = np.random.default_rng(5)
    x = np.linspace(0, 2 * np.pi, n, endpoint=False)
    X, Y, Z = np.meshgrid(x, x, x, indexing="ij")
    f = {"ux": np.sin(X) * np.cos(Y) * np.cos(Z), "uy": -np.cos(X) * np.sin(Y) * np.cos(Z), "uz": 0.1 * np.sin(2 * Z + X),
         "bx": np.sin(Y + Z), "by": np.cos(2 * X) * np.sin(Z), "bz": 0.3 * np.cos(X + Y)}
    f = {k: (v + 0.01 * rng.standard_normal(v.shape)).astype(np.float32) for k, v in f.items()}
    if phys == "hd":
        M.SetUpProblemIC(p, ux=f["ux"], uy=f["uy"], uz=f["uz"])
    elif phys == "mhd":
        M.SetUpProblemIC(p, **f)
    else:
        M.SetUpProblemIC(p, bx=f["bx"], by=f["by"], bz=f["bz"])
    M.stepforward(p, steps)
    sol = np.array(p.sol)
    p.close()
    return sol


if __name__ == "__main__":
    for phys in ("hd", "mhd", "emhd"):
        for stepper in ("RK4", "LSRK54"):
            a, b = run("0", phys, stepper), run("1", phys, stepper)
            print(f"spec2-vs-default {phys} {stepper} max abs diff {np.max(np.abs(a - b)):.3e} norm {np.linalg.norm(a):.3e}", flush=True)
