"""Small workload for compute-sanitizer (memcheck / racecheck / synccheck): every kernel family on tiny grids -- HD, MHD,
EMHD (both x-kernel forms), LSRK54, A99 driving, volume penalisation, divergence corrections, diagnostics, get/set.
    compute-sanitizer --tool memcheck python tools/sanitize_target.py
Sanitizer runs are 10-100x slower than native: grids stay at 32..64 points."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mhdflows_jl_b200 as M  # noqa: E402

rng = np.random.default_rng(3)


def fields(shape, n):
    return [rng.standard_normal(shape).astype(np.float32) * 0.1 for _ in range(n)]


def run(label, setup=None, steps=2, **kw):
    p = M.Problem(M.GPU(), **kw)
    f = fields(p._real_shape, 6)
    if kw.get("EMHD"):
        M.SetUpProblemIC(p, bx=f[3], by=f[4], bz=f[5])
    elif kw.get("B_field"):
        M.SetUpProblemIC(p, ux=f[0], uy=f[1], uz=f[2], bx=f[3], by=f[4], bz=f[5])
    else:
        M.SetUpProblemIC(p, ux=f[0], uy=f[1], uz=f[2])
    if setup:
        setup(p)
    M.stepforward(p, steps)
    p.calcN()
    p.energy(M.FRESH), p.energy(M.STALE), p.helicity(), M.spectralline(p, 0)
    p.get_real(0, M.STALE)
    print("sanitize-case", label, "ok", flush=True)
    p.close()


run("hd 32x64x32 rk4", nx=32, ny=64, nz=32, nu=1e-2, dt=1e-3)
run("mhd 64x32x32 rk4", nx=64, ny=32, nz=32, nu=1e-2, eta=1e-2, dt=1e-3, B_field=True)
run("mhd 32^3 lsrk54 hyper", nx=32, nu=1e-2, eta=1e-2, n_nu=2, dt=1e-3, B_field=True, stepper="LSRK54")
run("emhd 32^3 rk4", nx=32, dt=1e-4, B_field=True, EMHD=True)
os.environ["MHDF_EMHD2"] = "0"
run("emhd 32x32x64 rk4 register form of the x kernel", nx=32, ny=32, nz=64, dt=1e-4, B_field=True, EMHD=True)
os.environ["MHDF_EMHD2"] = "1"
uv, fn = M.GetA99vars_And_function(M.GPU(), 32, 32, 32)
run("mhd 32^3 a99 driving", setup=lambda p: M.SetUpFk(p, kf=2, P=1e-3), nx=32, nu=1e-2, eta=1e-2, dt=1e-3, B_field=True, calcF=fn, usr_vars=uv)


def vp_setup(p):
    p.params.χ = M.Cylindrical_Mask_Function(p.grid, R2=2.0)
    M.DivVCorrection(p)
    M.DivBCorrection(p)


run("mhd 32^3 volume penalisation + div corrections", setup=vp_setup, nx=32, nu=1e-2, eta=1e-2, dt=1e-3, B_field=True, VP_method=True)
run("mhd f64 32^3", nx=32, nu=1e-2, eta=1e-2, dt=1e-3, B_field=True, T=np.float64)


def ic_and_analysis(p):      # device-side DivFreeSpectraMap, ScaleDecomposition, VectorPotential
    M.SetUpRandomPhaseIC(p, seed_u=3, seed_b=4, k0=-5 / 6)
    M.ScaleDecomposition(p, "u", kf=[1, 4])
    M.VectorPotential(p, which=M.FRESH)


run("mhd 32x64x32 random-phase IC + on-device analysis", setup=ic_and_analysis, nx=32, ny=64, nz=32, nu=1e-2, eta=1e-2, dt=1e-3, B_field=True)
print("sanitize-target done")
