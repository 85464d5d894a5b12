"""GPU parity tests of the SURVEY 8f rows: A99 random driving (both reference implementations), DivVCorrection! /
DivBCorrection! and the volume-penalisation terms, through the C ABI against oracle/forcing_oracle.py.
(First green hardware run: round 2, profiles/r02_c1_pytest.log; the kernels are also checked on the CPU emulator,
tests/cpu_emu: Philox known answers, forcing of every mode, k_divclean.)"""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

F32_TOL = 1e-5
F64_TOL = 1e-12


@pytest.fixture(scope="module")
def M():
    import mhdflows_jl_b200 as M
    return M


@pytest.fixture(scope="module")
def O():
    from oracle import mhdflows_oracle as O
    return O


@pytest.fixture(scope="module")
def FO():
    from oracle import forcing_oracle as FO
    return FO


def _band_limited(g, x):
    """A real field whose spectrum lies strictly inside the band both implementations carry: dealias!() keeps the waves
    -n/3 .. n/3-1, so the unpaired wave -n/3 is dropped too (its Hermitian partner +n/3 is aliased: the reference keeps it in
    `vars` until the next dealias!, the library never stores it)."""
    h = g.dealias(g.rfft(x))
    h[:, g.ny - g.ny // 3, :] = 0
    h[g.nz - g.nz // 3, :, :] = 0
    return g.irfft(h)


def _forced_pair(M, O, FO, variant, T, dims=(32, 32, 32), stepper="RK4", seed=4242):
    nx, ny, nz = dims
    kw = dict(nx=nx, ny=ny, nz=nz, T=T, nu=2e-2, eta=3e-2, dt=4e-3, B_field=True, stepper=stepper)
    if variant == "host":
        op = O.Problem(calcF=FO.A99ForceDriving, **kw)
        op.vars.usr_vars = FO.A99Vars(op.grid)
        FO.SetUpFk(op, kf=3, P=2, sigma2=1)
        uv, fn = M.GetA99vars_And_function(M.GPU(), nx, ny, nz, T=T, seed=seed)
        gp = M.Problem(M.GPU(), calcF=fn, usr_vars=uv, **kw)
        M.SetUpFk(gp, kf=3, P=2, σ2=1)
    else:
        op = O.Problem(calcF=FO.A99ForceDriving_GPU, **kw)
        op.vars.usr_vars = FO.A99GPUVars(op.grid)
        FO.SetUpFk_GPU(op, kf=2.0, P=2e-3, sigma=1.5, b=0.8)      # the amplitude enters squared in this variant
        uv, fn, setup = M.A99GPU.GetA99vars_And_function(M.GPU(), nx, ny, nz, T=T, seed=seed)
        gp = M.Problem(M.GPU(), calcF=fn, usr_vars=uv, **kw)
        setup(gp, kf=2.0, P=2e-3, σ=1.5, b=0.8)
    op.vars.usr_vars.rng = FO.PhiloxField(seed, op.grid)
    u, b = O.random_phase_ic(op.grid, 21), O.random_phase_ic(op.grid, 22)
    O.SetUpProblemIC(op, *u, bx=b[0], by=b[1], bz=b[2])
    M.SetUpProblemIC(gp, ux=u[0], uy=u[1], uz=u[2], bx=b[0], by=b[1], bz=b[2])
    return op, gp


@pytest.mark.parametrize("variant", ["host", "gpu"])
@pytest.mark.parametrize("T,tol", [(np.float32, F32_TOL), (np.float64, F64_TOL)])
def test_a99_calcN_and_steps(M, O, FO, variant, T, tol):
    op, gp = _forced_pair(M, O, FO, variant, T, dims=(32, 16, 64))
    g = op.grid
    # one RHS evaluation (forcing call 0 on both sides)
    N = np.zeros_like(op.sol)
    op.calcN(N, op.sol.copy(), 0.0, op.clock, op.vars, op.params, g)
    Nd = gp.calcN()
    ref = g.dealias(N.copy())
    assert O.rel_l2(Nd, ref) < tol
    # the forcing is a visible part of N
    q = O.Problem(nx=g.nx, ny=g.ny, nz=g.nz, T=T, nu=2e-2, eta=3e-2, dt=4e-3, B_field=True)
    q.sol[...] = op.sol
    N0 = np.zeros_like(op.sol)
    q.calcN(N0, q.sol.copy(), 0.0, q.clock, q.vars, q.params, g)
    assert O.rel_l2(ref[:3], g.dealias(N0.copy())[:3]) > 1e-3
    assert gp.a99_calls() == 1
    # five steps: forcing calls 1..20
    for _ in range(5):
        O.stepforward(op)
    M.stepforward(gp, 5)
    assert gp.a99_calls() == 21 and op.vars.usr_vars.calls == 21
    assert O.rel_l2(gp.sol, g.dealias(op.sol.copy())) < tol
    gp.close()


def test_a99_lsrk54_and_retuned_amplitude(M, O, FO):
    op, gp = _forced_pair(M, O, FO, "host", np.float32, stepper="LSRK54")
    for _ in range(3):
        O.stepforward(op)
    M.stepforward(gp, 3)
    assert gp.a99_calls() == 15
    assert O.rel_l2(gp.sol, op.grid.dealias(op.sol.copy())) < F32_TOL
    # usr_vars.A / b are read on every forcing call by the reference: retune between steps
    op.vars.usr_vars.A = np.float32(2.5)
    op.vars.usr_vars.b = np.float32(0.6)
    gp.vars.usr_vars.A = np.float32(2.5)
    gp.vars.usr_vars.b = np.float32(0.6)
    for _ in range(2):
        O.stepforward(op)
    M.stepforward(gp, 2)
    assert gp.a99_calls() == 25
    assert O.rel_l2(gp.sol, op.grid.dealias(op.sol.copy())) < F32_TOL
    gp.close()


def test_a99_is_reproducible_and_hd_forcing_is_lost(M, O, FO):
    sols = []
    for _ in range(2):
        _, gp = _forced_pair(M, O, FO, "gpu", np.float32)
        M.stepforward(gp, 3)
        sols.append(gp.sol)
        gp.close()
    assert np.array_equal(sols[0], sols[1])                # same seed, same stream
    # HD: the forcing is added before the advection zeroes N (pgen.jl:176-178, HDSolver.jl:55)
    kw = dict(nx=32, T=np.float32, nu=2e-2, dt=4e-3)
    uv, fn = M.GetA99vars_And_function(M.GPU(), 32, 32, 32)
    forced, plain = M.Problem(M.GPU(), calcF=fn, usr_vars=uv, **kw), M.Problem(M.GPU(), **kw)
    M.SetUpFk(forced)
    u = O.random_phase_ic(O.Grid(32, T=np.float32), 3)
    for p in (forced, plain):
        M.SetUpProblemIC(p, ux=u[0], uy=u[1], uz=u[2])
        M.stepforward(p, 2)
    assert np.array_equal(forced.sol, plain.sol)
    forced.close()
    plain.close()


@pytest.mark.parametrize("T,tol", [(np.float32, F32_TOL), (np.float64, F64_TOL)])
def test_div_corrections(M, O, FO, T, tol):
    kw = dict(nx=32, ny=16, nz=64, T=T, nu=2e-2, eta=3e-2, dt=4e-3, B_field=True)
    op, gp = O.Problem(**kw), M.Problem(M.GPU(), **kw)
    g = op.grid
    rng = np.random.default_rng(8)
    # band-limited but not solenoidal fields (band-limited: the library stores the dealiased band only)
    f = [_band_limited(g, rng.standard_normal((64, 16, 32)).astype(T)) for _ in range(6)]
    O.SetUpProblemIC(op, *f[:3], bx=f[3], by=f[4], bz=f[5])
    M.SetUpProblemIC(gp, ux=f[0], uy=f[1], uz=f[2], bx=f[3], by=f[4], bz=f[5])
    FO.DivBCorrection(op)
    M.DivBCorrection(gp)
    assert O.rel_l2(gp.sol, g.dealias(op.sol.copy())) < tol
    assert O.rel_l2(gp.vars.bx, op.vars.bx) < 10 * tol and O.rel_l2(gp.vars.ux, op.vars.ux) < 10 * tol
    FO.DivVCorrection(op)
    M.DivVCorrection(gp)
    sol = gp.sol
    assert O.rel_l2(sol, g.dealias(op.sol.copy())) < tol
    for base in (0, 3):
        div = g.kr * sol[base] + g.l * sol[base + 1] + g.m * sol[base + 2]
        assert np.linalg.norm(div.ravel()) / np.linalg.norm(sol[base:base + 3].ravel()) < (1e-5 if T is np.float32 else 1e-13)
    # dashboard energies and CFL maxima follow the refreshed vars
    ke, me = gp.energy(M.STALE)
    dV = float(T(g.dx)) * float(T(g.dy)) * float(T(g.dz))
    ke_ref = sum(float(np.sum(getattr(op.vars, n).astype(np.float64) ** 2)) for n in ("ux", "uy", "uz")) * dV
    me_ref = sum(float(np.sum(getattr(op.vars, n).astype(np.float64) ** 2)) for n in ("bx", "by", "bz")) * dV
    assert abs(ke - ke_ref) < 1e-4 * ke_ref and abs(me - me_ref) < 1e-4 * me_ref
    mx, _ = gp.stale_stats()
    assert abs(mx[4] - float(np.max(op.vars.by.astype(np.float64) ** 2))) < 1e-4 * mx[4]
    # the corrections stay consistent with stepping afterwards
    for _ in range(3):
        O.stepforward(op)
    M.stepforward(gp, 3)
    assert O.rel_l2(gp.sol, g.dealias(op.sol.copy())) < tol
    with pytest.raises(M.MHDFlowsError):
        gp.div_correction(2)
    gp.close()
    hd = M.Problem(M.GPU(), nx=16, T=T)
    with pytest.raises(M.MHDFlowsError):
        M.DivBCorrection(hd)
    hd.close()


def test_div_b_correction_emhd(M, O, FO):
    kw = dict(nx=32, T=np.float32, B_field=True, EMHD=True, dt=2e-4)
    op, gp = O.Problem(**kw), M.Problem(M.GPU(), **kw)
    g = op.grid
    rng = np.random.default_rng(9)
    f = [_band_limited(g, rng.standard_normal((32, 32, 32)).astype(np.float32)) for _ in range(3)]
    O.SetUpProblemIC(op, bx=f[0], by=f[1], bz=f[2])
    M.SetUpProblemIC(gp, bx=f[0], by=f[1], bz=f[2])
    FO.DivBCorrection(op)
    M.DivBCorrection(gp)
    assert O.rel_l2(gp.sol, g.dealias(op.sol.copy())) < F32_TOL
    # EMHD's (B.grad)A term reads the stale real-space vars.b*: they were refreshed by the correction
    for _ in range(3):
        O.stepforward(op)
    M.stepforward(gp, 3)
    assert O.rel_l2(gp.sol, g.dealias(op.sol.copy())) < F32_TOL
    gp.close()


def _vp_pair(M, O, B, T, dims=(32, 16, 64)):
    nx, ny, nz = dims
    kw = dict(nx=nx, ny=ny, nz=nz, T=T, nu=2e-2, dt=2e-3, VP_method=True)
    if B:
        kw.update(eta=3e-2, B_field=True)
    op, gp = O.Problem(**kw), M.Problem(M.GPU(), **kw)
    g = op.grid
    u, b = O.random_phase_ic(g, 31), O.random_phase_ic(g, 32)
    # a cylinder of solid around the z axis, rotating like a rigid body, with a uniform field frozen into it
    X, Y = g.x.astype(np.float64).reshape(1, 1, -1), g.y.astype(np.float64).reshape(1, -1, 1)
    chi = (((X ** 2 + Y ** 2) < 1.5) * np.ones((nz, ny, nx))).astype(T)
    U0 = [(-0.3 * Y * chi).astype(T), (0.3 * X * chi).astype(T), np.zeros((nz, ny, nx), T)]
    B0 = [np.zeros((nz, ny, nx), T), np.zeros((nz, ny, nx), T), (0.2 * chi).astype(T)]
    op.params.vp.chi[...] = chi
    op.params.vp.U0x[...], op.params.vp.U0y[...], op.params.vp.U0z[...] = U0
    gp.params.χ = chi
    if B:
        op.params.vp.B0x[...], op.params.vp.B0y[...], op.params.vp.B0z[...] = B0
        O.SetUpProblemIC(op, *u, bx=b[0], by=b[1], bz=b[2])
        M.SetUpProblemIC(gp, ux=u[0], uy=u[1], uz=u[2], bx=b[0], by=b[1], bz=b[2],
                         U0x=U0[0], U0y=U0[1], U0z=U0[2], B0x=B0[0], B0y=B0[1], B0z=B0[2])     # Python cannot spell U₀x
    else:
        O.SetUpProblemIC(op, *u)
        M.SetUpProblemIC(gp, ux=u[0], uy=u[1], uz=u[2], U0x=U0[0], U0y=U0[1], U0z=U0[2])
    return op, gp


@pytest.mark.parametrize("B", [False, True])
@pytest.mark.parametrize("T,tol", [(np.float32, F32_TOL), (np.float64, F64_TOL)])
def test_volume_penalisation(M, O, B, T, tol):
    """Problem(...; VP_method = true): the penalisation terms of VPSolver.jl:21-59 inside every RHS evaluation."""
    op, gp = _vp_pair(M, O, B, T)
    g = op.grid
    N = np.zeros_like(op.sol)
    op.calcN(N, op.sol.copy(), 0.0, op.clock, op.vars, op.params, g)
    assert O.rel_l2(gp.calcN(), g.dealias(N.copy())) < tol
    q = O.Problem(nx=g.nx, ny=g.ny, nz=g.nz, T=T, nu=2e-2, dt=2e-3, **(dict(eta=3e-2, B_field=True) if B else {}))
    q.sol[...] = op.sol
    N0 = np.zeros_like(op.sol)
    q.calcN(N0, q.sol.copy(), 0.0, q.clock, q.vars, q.params, g)
    assert O.rel_l2(g.dealias(N.copy()), g.dealias(N0.copy())) > 1e-2          # the penalisation is a visible part of N
    for _ in range(5):
        O.stepforward(op)
    M.stepforward(gp, 5)
    assert O.rel_l2(gp.sol, g.dealias(op.sol.copy())) < tol
    gp.close()


def test_volume_penalisation_time_integrator(M, O):
    """TimeIntegrator! with flag.vp: DivVCorrection! / DivBCorrection! before the loop and after every step
    (integrator.jl:85-88, 118-122); eta follows clock.dt."""
    op, gp = _vp_pair(M, O, True, np.float32, dims=(32, 32, 32))
    O.TimeIntegrator(op, 1e9, 3, usr_dt=1.5e-3)
    M.TimeIntegrator(gp, 1e9, 3, usr_dt=1.5e-3)
    g = op.grid
    assert gp.clock.step == op.clock.step == 4
    assert O.rel_l2(gp.sol, g.dealias(op.sol.copy())) < F32_TOL
    sol = gp.sol
    for base in (0, 3):
        div = g.kr * sol[base] + g.l * sol[base + 1] + g.m * sol[base + 2]
        assert np.linalg.norm(div.ravel()) / np.linalg.norm(sol[base:base + 3].ravel()) < 1e-5
    # Documented deviation: after a VP step the reference's vars.* still carry the aliased-band content of the
    # full sol, the library's do not.  Its effect on what getCFL! computes from those vars is bounded here.
    dt_ref, dt_gpu = O.getCFL(op, math.inf, Coef=0.25), M.getCFL(gp, math.inf, Coef=0.25)
    assert abs(dt_gpu - dt_ref) / dt_ref < 2e-2, (dt_gpu, dt_ref)
    gp.close()
    with pytest.raises(ValueError):
        M.Problem(M.GPU(), nx=16, B_field=True, EMHD=True, VP_method=True)


def _nd_pair(M, O, FO, T, dims=(32, 32, 32), stepper="RK4", B_field=True):
    nx, ny, nz = dims
    kw = dict(nx=nx, ny=ny, nz=nz, T=T, nu=2e-2, dt=4e-3, stepper=stepper)
    if B_field:
        kw.update(eta=3e-2, B_field=True)
    op = O.Problem(calcF=FO.NDForceDriving, **kw)
    op.vars.usr_vars = FO.NDVars(op.grid)
    uv, fn = M.GetNDvars_And_function(M.GPU(), nx, ny, nz, T=T)
    gp = M.Problem(M.GPU(), calcF=fn, usr_vars=uv, **kw)
    g = op.grid
    x, y, z = g.x.astype(np.float64).reshape(1, 1, -1), g.y.astype(np.float64).reshape(1, -1, 1), g.z.astype(np.float64).reshape(-1, 1, 1)
    fx = (1.0 + 0.5 * np.sin(x) * np.cos(2 * y) + 0 * z).astype(T)
    fy = (0.7 * np.cos(x + z) + 0 * y).astype(T)
    fz = (np.sin(y) * np.sin(z) - 0.2 + 0 * x).astype(T)
    FO.SetUpND(op, 0.35, fx, fy, fz)
    M.SetUpND(gp, 0.35, fx, fy, fz)
    u, b = O.random_phase_ic(g, 41), O.random_phase_ic(g, 42)
    if B_field:
        O.SetUpProblemIC(op, *u, bx=b[0], by=b[1], bz=b[2])
        M.SetUpProblemIC(gp, ux=u[0], uy=u[1], uz=u[2], bx=b[0], by=b[1], bz=b[2])
    else:
        O.SetUpProblemIC(op, *u)
        M.SetUpProblemIC(gp, ux=u[0], uy=u[1], uz=u[2])
    return op, gp


@pytest.mark.parametrize("T,tol,stepper", [(np.float32, F32_TOL, "RK4"), (np.float64, F64_TOL, "RK4"), (np.float32, F32_TOL, "LSRK54")])
def test_negative_damping_forcing(M, O, FO, T, tol, stepper):
    """NDForceDriving! + SetUpND! (pgen/NegativeDamping.jl:14-45): N_ui += P rfft(f_i u_i) / (sum |u_i^2 f_i| dV) on every RHS
    evaluation of an MHD problem, formed inside the fused x kernel and normalised by its own reduction."""
    op, gp = _nd_pair(M, O, FO, T, dims=(32, 16, 64), stepper=stepper)
    g = op.grid
    N = np.zeros_like(op.sol)
    op.calcN(N, op.sol.copy(), 0.0, op.clock, op.vars, op.params, g)
    ref = g.dealias(N.copy())
    assert O.rel_l2(gp.calcN(), ref) < tol
    q = O.Problem(nx=g.nx, ny=g.ny, nz=g.nz, T=T, nu=2e-2, eta=3e-2, dt=4e-3, B_field=True)      # the forcing really acts
    q.sol[...] = op.sol
    N0 = np.zeros_like(op.sol)
    q.calcN(N0, q.sol.copy(), 0.0, q.clock, q.vars, q.params, g)
    assert O.rel_l2(ref[:3], g.dealias(N0.copy())[:3]) > 1e-3 and O.rel_l2(ref[3:], g.dealias(N0.copy())[3:]) == 0.0
    for _ in range(3):
        O.stepforward(op)
    M.stepforward(gp, 3)
    assert O.rel_l2(gp.sol, g.dealias(op.sol.copy())) < tol
    gp.close()


def test_negative_damping_is_lost_in_hd_and_refused_with_vp(M, O, FO):
    op, gp = _nd_pair(M, O, FO, np.float32, dims=(32, 32, 32), B_field=False)     # HDcalcN! clobbers the forcing (pgen.jl:176-178)
    O.stepforward(op)
    M.stepforward(gp)
    assert O.rel_l2(gp.sol, op.grid.dealias(op.sol.copy())) < F32_TOL
    gp.close()
    uv, fn = M.GetNDvars_And_function(M.GPU(), 32, 32, 32)
    with pytest.raises(NotImplementedError):
        M.Problem(M.GPU(), nx=32, B_field=True, calcF=fn, usr_vars=uv, VP_method=True)
    with pytest.raises(ValueError):
        M.Problem(M.GPU(), nx=32, B_field=True, calcF=fn)


def _structure_function_check(M, O, T, tol, dims):
    """CF / SFC / SF₂1D (utils/TurbStatTool.jl:67, 72, 90-120) of the state's fields on the device against the literal restatement
    applied to the reference's vars.* (the stale real-space fields a script would pass)."""
    nx, ny, nz = dims
    kw = dict(nx=nx, ny=ny, nz=nz, T=T, nu=2e-2, eta=3e-2, dt=4e-3, B_field=True)
    op, gp = O.Problem(**kw), M.Problem(M.GPU(), **kw)
    u, b = O.random_phase_ic(op.grid, 31), O.random_phase_ic(op.grid, 32)
    O.SetUpProblemIC(op, *u, bx=b[0], by=b[1], bz=b[2])
    M.SetUpProblemIC(gp, ux=u[0], uy=u[1], uz=u[2], bx=b[0], by=b[1], bz=b[2])
    O.stepforward(op)
    M.stepforward(gp)
    for grp, v in (("u", (op.vars.ux, op.vars.uy, op.vars.uz)), ("b", (op.vars.bx, op.vars.by, op.vars.bz))):
        for got, ref in zip(M.CF(gp, grp), (O.CF(x) for x in v)):
            assert np.linalg.norm(ref) > 0 and O.rel_l2(got, ref) < 3 * tol      # quadratic in the field: twice its relative error
        for got, ref in zip(M.SFC(gp, grp), (O.SFC(x) for x in v)):
            assert O.rel_l2(got, ref) < 3 * tol
    got, ref = M.SF2_1D(gp, "u"), O.SF2_1D(op.vars.ux, op.vars.uz, op.vars.uy)
    ok = np.isfinite(ref)
    assert got.shape == ref.shape and np.array_equal(np.isfinite(got), ok) and ok.sum() > 3
    assert np.linalg.norm(got[ok] - ref[ok]) < 3 * tol * np.linalg.norm(ref[ok])
    gp.close()


@pytest.mark.parametrize("T,tol", [(np.float32, F32_TOL), (np.float64, F64_TOL)])
def test_structure_functions_on_device(M, O, T, tol):
    _structure_function_check(M, O, T, tol, (32, 16, 16))


def _closure_forcing_check(M, O, T, tol, dims, stepper, B_field=True, steps=2):
    """An arbitrary calcF! closure (pgen.jl:231-234) -- time dependent and reading sol -- through the host-callback path of the
    library against the same function handed to the restatement: stage times, the `sol` argument and the upload must all agree."""
    nx, ny, nz = dims
    g0 = O.Grid(nx, ny, nz, T=T)
    X, Y, Z = (g0.x.astype(np.float64).reshape(1, 1, -1), g0.y.astype(np.float64).reshape(1, -1, 1), g0.z.astype(np.float64).reshape(-1, 1, 1))
    fxh = g0.rfft((0.5 * np.sin(2 * X) * np.cos(2 * Y) * np.cos(Z)).astype(T))
    fyh = g0.rfft((-0.5 * np.cos(2 * X) * np.sin(2 * Y) * np.cos(Z)).astype(T))
    seen = {0: [], 1: []}

    def make(off):           # the mirror's params.*_ind are the reference's 1-based numbers, the restatement's are 0-based
        def calcF(N, sol, t, clock, vars, params, grid):
            seen[off].append(float(t))
            N[params.ux_ind - off] += T(math.cos(40 * t)) * fxh - T(0.3) * sol[params.uy_ind - off]
            N[params.uy_ind - off] += T(1 + 10 * t) * fyh
            if B_field:
                N[params.bz_ind - off] += T(0.05) * sol[params.bx_ind - off]
        return calcF

    kw = dict(nx=nx, ny=ny, nz=nz, T=T, stepper=stepper, nu=2e-2, dt=4e-3, B_field=B_field)
    if B_field:
        kw["eta"] = 3e-2
    op, gp = O.Problem(calcF=make(0), **kw), M.Problem(M.GPU(), calcF=make(1), **kw)
    u, b = O.random_phase_ic(op.grid, 41), O.random_phase_ic(op.grid, 42)
    if B_field:
        O.SetUpProblemIC(op, *u, bx=b[0], by=b[1], bz=b[2])
        M.SetUpProblemIC(gp, ux=u[0], uy=u[1], uz=u[2], bx=b[0], by=b[1], bz=b[2])
    else:
        O.SetUpProblemIC(op, *u)
        M.SetUpProblemIC(gp, ux=u[0], uy=u[1], uz=u[2])
    for _ in range(steps):
        O.stepforward(op)
    M.stepforward(gp, steps)
    assert len(seen[1]) == len(seen[0]) == steps * (4 if stepper == "RK4" else 5)
    assert np.allclose(seen[1], seen[0], rtol=1e-6 if T is np.float32 else 1e-14, atol=0)      # FourierFlows' stage times
    assert O.rel_l2(gp.sol, op.grid.dealias(op.sol.copy())) < tol
    gp.close()
    return op


@pytest.mark.parametrize("T,tol,stepper", [(np.float32, F32_TOL, "RK4"), (np.float64, F64_TOL, "LSRK54")])
def test_arbitrary_calcF_closure_through_the_host_callback(M, O, T, tol, stepper):
    op = _closure_forcing_check(M, O, T, tol, (32, 16, 32), stepper)
    q = O.Problem(nx=32, ny=16, nz=32, T=T, stepper=stepper, nu=2e-2, eta=3e-2, dt=4e-3, B_field=True)     # the forcing really acts
    u, b = O.random_phase_ic(q.grid, 41), O.random_phase_ic(q.grid, 42)
    O.SetUpProblemIC(q, *u, bx=b[0], by=b[1], bz=b[2])
    for _ in range(2):
        O.stepforward(q)
    assert O.rel_l2(q.grid.dealias(q.sol.copy()), op.grid.dealias(op.sol.copy())) > 1e-4


def test_calcF_closure_lost_in_hd_and_exceptions_propagate(M, O):
    _closure_forcing_check(M, O, np.float32, F32_TOL, (32, 32, 32), "RK4", B_field=False, steps=1)   # HDcalcN! clobbers it (pgen.jl:176-178)

    def bad(N, sol, t, clock, vars, params, grid):
        raise KeyError("boom")
    p = M.Problem(M.GPU(), nx=32, B_field=True, dt=1e-3, calcF=bad)
    p.set_real("ux", np.ones((32, 32, 32), dtype=np.float32))
    with pytest.raises(KeyError):
        M.stepforward(p)
    p.close()
    with pytest.raises(NotImplementedError):
        M.Problem(M.GPU(), nx=32, B_field=True, EMHD=True, stepper="HM89", calcF=bad)


def _random_phase_case(M, O, FO, T, tol, dims, nranks_note=""):
    """mhdf_set_random_phase against oracle.DivFreeSpectraMap with the device's Philox phases injected (IC.jl:130-179)."""
    nx, ny, nz = dims
    L3 = dict(Lx=2 * np.pi, Ly=3.0, Lz=5.0)
    gp = M.Problem(M.GPU(), nx=nx, ny=ny, nz=nz, T=T, nu=1e-2, eta=1e-2, dt=1e-3, B_field=True, **L3)
    M.SetUpRandomPhaseIC(gp, seed_u=1234, seed_b=(5 << 32) + 678, k0=-5 / 6, P=2.0, k_peak=1.5)
    g = O.Grid(nx, ny, nz, L3["Lx"], L3["Ly"], L3["Lz"], T)
    worst = 0.0
    for names, seed in ((("ux", "uy", "uz"), 1234), (("bx", "by", "bz"), (5 << 32) + 678)):
        theta = FO.PhiloxField(seed, g).uniforms(M.DFSM_CALL)[0]
        ref = O.DivFreeSpectraMap(g, theta, k_peak=1.5, P=2.0, k0=-5 / 6)
        for nm, r in zip(names, ref):
            assert np.linalg.norm(r) > 0
            worst = max(worst, O.rel_l2(gp.get_real(nm, M.FRESH), r), O.rel_l2(gp.get_real(nm, M.STALE), r))
    assert worst < tol, worst
    # SetUpProblemIC! semantics: sol = rfft(F), vars.* = F  ->  the stale statistics are those of F
    op = O.Problem(nx=nx, ny=ny, nz=nz, T=T, nu=1e-2, eta=1e-2, dt=1e-3, B_field=True, **L3)
    u = O.DivFreeSpectraMap(g, FO.PhiloxField(1234, g).uniforms(M.DFSM_CALL)[0], k_peak=1.5, P=2.0, k0=-5 / 6)
    b = O.DivFreeSpectraMap(g, FO.PhiloxField((5 << 32) + 678, g).uniforms(M.DFSM_CALL)[0], k_peak=1.5, P=2.0, k0=-5 / 6)
    O.SetUpProblemIC(op, *u, bx=b[0], by=b[1], bz=b[2])
    ke, me = gp.energy(M.STALE)
    ko, mo = O.ProbDiagnostic(op, rounded=False)
    assert abs(ke - ko) < 1e-4 * abs(ko) and abs(me - mo) < 1e-4 * abs(mo)
    assert O.rel_l2(gp.sol, op.grid.dealias(op.sol.copy())) < tol
    O.stepforward(op)
    M.stepforward(gp)
    assert O.rel_l2(gp.sol, op.grid.dealias(op.sol.copy())) < tol
    gp.close()
    return worst


@pytest.mark.parametrize("T,tol,dims", [(np.float32, F32_TOL, (32, 64, 16)), (np.float64, F64_TOL, (32, 16, 32))])
def test_random_phase_ic_on_device(M, O, FO, T, tol, dims):
    """DivFreeSpectraMap + SetUpProblemIC! on the device (mhdf_set_random_phase) against the restatement of IC.jl:130-179 fed with the
    device's Philox phases, on a non-cubic box with three different side lengths."""
    _random_phase_case(M, O, FO, T, tol, dims)
