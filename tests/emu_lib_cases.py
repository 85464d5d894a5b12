"""End-to-end cases run against the CPU-EMULATED build of the library (tests/cpu_emu/cuda_host_emu.h): the real api.cu +
solver.cuh orchestration and the real kernel sources, compiled as plain C++, on 16-point grids.  Executed by
tests/test_emulated_library.py in a subprocess with MHDF_LIB pointing at the emulated library; prints one PASS / FAIL line
per case.  This is how features written without GPU access (A99 driving, volume penalisation, divergence corrections, the
second EMHD kernel form, HDF5 dumps) get their host-side logic exercised before their first hardware run -- the emulated
library is test infrastructure, never a fallback of the product (mhdflows_jl_b200 only ever loads it through MHDF_LIB).

Not collected by pytest (no test_ prefix)."""
import os
import sys
import tempfile
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mhdflows_jl_b200 as M  # noqa: E402
from oracle import forcing_oracle as FO  # noqa: E402
from oracle import mhdflows_oracle as O  # noqa: E402
from tests.test_gpu_parity import _hm89_check, _pair  # noqa: E402
from tests.test_gpu_zforcing import (_closure_forcing_check, _forced_pair, _nd_pair, _random_phase_case,  # noqa: E402
                                     _structure_function_check, _vp_pair)

F32_TOL, F64_TOL = 1e-5, 1e-12
DIMS = (16, 16, 32)
CASES = []


def case(fn):
    CASES.append(fn)
    return fn


def _band_limited(g, x):
    """A real field whose spectrum lies strictly inside the band both implementations carry: dealias!() keeps the waves
    -n/3 .. n/3-1, so the unpaired wave -n/3 is dropped too (its Hermitian partner +n/3 is aliased: the reference keeps it in
    `vars` until the next dealias!, the library never stores it)."""
    h = g.dealias(g.rfft(x))
    h[:, g.ny - g.ny // 3, :] = 0
    h[g.nz - g.nz // 3, :, :] = 0
    return g.irfft(h)


def _dealiased(op):
    return op.grid.dealias(op.sol.copy())


@case
def smoke_mhd_rk4_against_the_oracle():
    kw = dict(nx=16, T=np.float32, nu=1e-2, eta=1e-2, dt=5e-3, B_field=True)
    op, gp = O.Problem(**kw), M.Problem(M.GPU(), **kw)
    ic = O.taylor_green_ic(op.grid)
    O.SetUpProblemIC(op, *ic[:3], bx=ic[3], by=ic[4], bz=ic[5])
    M.SetUpProblemIC(gp, ux=ic[0], uy=ic[1], uz=ic[2], bx=ic[3], by=ic[4], bz=ic[5])
    O.stepforward(op)
    M.stepforward(gp)
    assert O.rel_l2(gp.sol, _dealiased(op)) < F32_TOL
    gp.close()


@case
def random_phase_ic_on_device_matches_the_oracle():
    _random_phase_case(M, O, FO, np.float32, F32_TOL, (16, 32, 16))
    _random_phase_case(M, O, FO, np.float64, F64_TOL, (16, 16, 16))


@case
def on_device_scale_decomposition_and_vector_potential():
    """mhdf_scale_decomposition / mhdf_vector_potential against the restatements of MHDAnalysis.jl:54-82, 129-174."""
    for T, tol in ((np.float32, F32_TOL), (np.float64, F64_TOL)):
        kw = dict(nx=16, ny=32, nz=16, Ly=3.0, T=T, nu=1e-2, eta=1e-2, dt=2e-3, B_field=True)
        op, gp = O.Problem(**kw), M.Problem(M.GPU(), **kw)
        g = op.grid
        u, b = O.random_phase_ic(g, 5), O.random_phase_ic(g, 6)
        O.SetUpProblemIC(op, *u, bx=b[0], by=b[1], bz=b[2])
        M.SetUpProblemIC(gp, ux=u[0], uy=u[1], uz=u[2], bx=b[0], by=b[1], bz=b[2])
        O.stepforward(op)
        M.stepforward(gp)
        vb = (op.vars.bx, op.vars.by, op.vars.bz)          # the stale vars: what a user script passes to the analysis functions
        vu = (op.vars.ux, op.vars.uy, op.vars.uz)
        # vectors are compared as a whole: a_z of these fields is 1e-3 of |a| (cancellation), its own relative error means nothing
        assert O.rel_l2(np.stack(M.VectorPotential(gp)), np.stack(O.VectorPotential(*vb, g))) < tol
        for grp, v in (("b", vb), ("u", vu)):
            for got, ref in zip(M.ScaleDecomposition(gp, grp, kf=[2, 4.5]), O.ScaleDecomposition(*v, g, kf=[2, 4.5])):
                assert np.linalg.norm(ref) > 0 and O.rel_l2(got, ref) < tol
        fresh = [g.irfft(g.dealias(op.sol[3 + i].copy())) for i in range(3)]
        assert O.rel_l2(np.stack(M.VectorPotential(gp, which=M.FRESH)), np.stack(O.VectorPotential(*fresh, g))) < tol
        gp.close()


@case
def on_device_structure_functions():
    """mhdf_correlation + the host-side SFC / SF_2 1D of the mirror against the restatements of TurbStatTool.jl:67, 72, 90-120."""
    _structure_function_check(M, O, np.float64, F64_TOL, (16, 16, 16))
    _structure_function_check(M, O, np.float32, F32_TOL, (16, 32, 16))


@case
def closure_forcing_through_the_host_callback():
    """Arbitrary calcF! closures: mhdf_set_forcing_callback / MHDF_STAGE / mhdf_set_forcing_spectral behind Problem(calcF = f)."""
    _closure_forcing_check(M, O, np.float64, F64_TOL, (16, 16, 16), "RK4")
    _closure_forcing_check(M, O, np.float32, F32_TOL, (16, 16, 32), "LSRK54", steps=1)
    _closure_forcing_check(M, O, np.float32, F32_TOL, (16, 16, 16), "RK4", B_field=False, steps=1)


@case
def negative_damping_forcing_calcN_and_steps():
    """NDForceDriving! (pgen/NegativeDamping.jl): products in the x kernel, normalisation by its reduction, added in the spectral kernel."""
    for T, tol, stepper in ((np.float32, F32_TOL, "RK4"), (np.float64, F64_TOL, "LSRK54")):
        op, gp = _nd_pair(M, O, FO, T, dims=DIMS, stepper=stepper)
        g = op.grid
        N = np.zeros_like(op.sol)
        op.calcN(N, op.sol.copy(), 0.0, op.clock, op.vars, op.params, g)
        ref = g.dealias(N.copy())
        err = O.rel_l2(gp.calcN(), ref)
        assert err < tol, err
        q = O.Problem(nx=g.nx, ny=g.ny, nz=g.nz, T=T, nu=2e-2, eta=3e-2, dt=4e-3, B_field=True)
        q.sol[...] = op.sol
        N0 = np.zeros_like(op.sol)
        q.calcN(N0, q.sol.copy(), 0.0, q.clock, q.vars, q.params, g)
        assert O.rel_l2(ref[:3], g.dealias(N0.copy())[:3]) > 1e-3
        for _ in range(2):
            O.stepforward(op)
        M.stepforward(gp, 2)
        assert O.rel_l2(gp.sol, _dealiased(op)) < tol
        gp.close()
    op, gp = _nd_pair(M, O, FO, np.float32, dims=(16, 16, 16), B_field=False)      # lost in HD, like the reference
    O.stepforward(op)
    M.stepforward(gp)
    assert O.rel_l2(gp.sol, _dealiased(op)) < F32_TOL
    gp.close()


@case
def a99_host_variant_calcN_steps_counter():
    op, gp = _forced_pair(M, O, FO, "host", np.float32, dims=DIMS)
    g = op.grid
    N = np.zeros_like(op.sol)
    op.calcN(N, op.sol.copy(), 0.0, op.clock, op.vars, op.params, g)
    Nd = gp.calcN()
    ref = g.dealias(N.copy())
    assert O.rel_l2(Nd, ref) < F32_TOL, O.rel_l2(Nd, ref)
    q = O.Problem(nx=g.nx, ny=g.ny, nz=g.nz, T=np.float32, nu=2e-2, eta=3e-2, dt=4e-3, B_field=True)
    q.sol[...] = op.sol
    N0 = np.zeros_like(op.sol)
    q.calcN(N0, q.sol.copy(), 0.0, q.clock, q.vars, q.params, g)
    assert O.rel_l2(ref[:3], g.dealias(N0.copy())[:3]) > 1e-3
    assert gp.a99_calls() == 1
    O.stepforward(op)
    M.stepforward(gp)
    assert gp.a99_calls() == 5 and op.vars.usr_vars.calls == 5
    assert O.rel_l2(gp.sol, _dealiased(op)) < F32_TOL
    gp.close()


@case
def a99_gpu_variant_calcN_and_step():
    op, gp = _forced_pair(M, O, FO, "gpu", np.float32, dims=DIMS)
    g = op.grid
    N = np.zeros_like(op.sol)
    op.calcN(N, op.sol.copy(), 0.0, op.clock, op.vars, op.params, g)
    assert O.rel_l2(gp.calcN(), g.dealias(N.copy())) < F32_TOL
    O.stepforward(op)
    M.stepforward(gp)
    assert O.rel_l2(gp.sol, _dealiased(op)) < F32_TOL
    gp.close()


@case
def a99_float64_both_variants_calcN():
    for variant in ("host", "gpu"):
        op, gp = _forced_pair(M, O, FO, variant, np.float64, dims=(16, 16, 16))
        g = op.grid
        N = np.zeros_like(op.sol)
        op.calcN(N, op.sol.copy(), 0.0, op.clock, op.vars, op.params, g)
        err = O.rel_l2(gp.calcN(), g.dealias(N.copy()))
        assert err < F64_TOL, (variant, err)
        gp.close()


@case
def a99_lsrk54_and_retuned_amplitude():
    op, gp = _forced_pair(M, O, FO, "host", np.float32, dims=(16, 16, 16), stepper="LSRK54")
    O.stepforward(op)
    M.stepforward(gp)
    assert gp.a99_calls() == 5
    op.vars.usr_vars.A = np.float32(2.5)
    op.vars.usr_vars.b = np.float32(0.6)
    gp.vars.usr_vars.A = np.float32(2.5)
    gp.vars.usr_vars.b = np.float32(0.6)
    O.stepforward(op)
    M.stepforward(gp)
    assert gp.a99_calls() == 10
    assert O.rel_l2(gp.sol, _dealiased(op)) < F32_TOL
    gp.close()


@case
def a99_reproducible_and_lost_in_hd():
    sols = []
    for _ in range(2):
        _, gp = _forced_pair(M, O, FO, "gpu", np.float32, dims=(16, 16, 16))
        M.stepforward(gp)
        sols.append(gp.sol)
        gp.close()
    assert np.array_equal(sols[0], sols[1])
    kw = dict(nx=16, T=np.float32, nu=2e-2, dt=4e-3)
    uv, fn = M.GetA99vars_And_function(M.GPU(), 16, 16, 16)
    forced, plain = M.Problem(M.GPU(), calcF=fn, usr_vars=uv, **kw), M.Problem(M.GPU(), **kw)
    M.SetUpFk(forced)
    u = O.random_phase_ic(O.Grid(16, T=np.float32), 3)
    for p in (forced, plain):
        M.SetUpProblemIC(p, ux=u[0], uy=u[1], uz=u[2])
        M.stepforward(p)
    assert np.array_equal(forced.sol, plain.sol)
    forced.close()
    plain.close()


def _div_case(T, tol):
    nx, ny, nz = DIMS
    kw = dict(nx=nx, ny=ny, nz=nz, T=T, nu=2e-2, eta=3e-2, dt=4e-3, B_field=True)
    op, gp = O.Problem(**kw), M.Problem(M.GPU(), **kw)
    g = op.grid
    rng = np.random.default_rng(8)
    f = [_band_limited(g, rng.standard_normal((nz, ny, nx)).astype(T)) for _ in range(6)]
    O.SetUpProblemIC(op, *f[:3], bx=f[3], by=f[4], bz=f[5])
    M.SetUpProblemIC(gp, ux=f[0], uy=f[1], uz=f[2], bx=f[3], by=f[4], bz=f[5])
    FO.DivBCorrection(op)
    M.DivBCorrection(gp)
    assert O.rel_l2(gp.sol, _dealiased(op)) < tol
    assert O.rel_l2(gp.vars.bx, op.vars.bx) < 10 * tol and O.rel_l2(gp.vars.ux, op.vars.ux) < 10 * tol
    FO.DivVCorrection(op)
    M.DivVCorrection(gp)
    sol = gp.sol
    assert O.rel_l2(sol, _dealiased(op)) < tol
    for base in (0, 3):
        div = g.kr * sol[base] + g.l * sol[base + 1] + g.m * sol[base + 2]
        assert np.linalg.norm(div.ravel()) / np.linalg.norm(sol[base:base + 3].ravel()) < (1e-5 if T is np.float32 else 1e-13)
    ke, me = gp.energy(M.STALE)
    dV = float(T(g.dx)) * float(T(g.dy)) * float(T(g.dz))
    ke_ref = sum(float(np.sum(getattr(op.vars, n).astype(np.float64) ** 2)) for n in ("ux", "uy", "uz")) * dV
    me_ref = sum(float(np.sum(getattr(op.vars, n).astype(np.float64) ** 2)) for n in ("bx", "by", "bz")) * dV
    assert abs(ke - ke_ref) < 1e-4 * ke_ref and abs(me - me_ref) < 1e-4 * me_ref
    mx, _ = gp.stale_stats()
    assert abs(mx[4] - float(np.max(op.vars.by.astype(np.float64) ** 2))) < 1e-4 * mx[4]
    O.stepforward(op)
    M.stepforward(gp)
    assert O.rel_l2(gp.sol, _dealiased(op)) < tol
    # after a step the stale view is a different register: the correction must reach it too
    FO.DivBCorrection(op)
    M.DivBCorrection(gp)
    assert O.rel_l2(gp.sol, _dealiased(op)) < tol
    assert O.rel_l2(gp.get_real("bz", M.STALE), g.irfft(_dealiased(op)[5])) < 10 * tol
    for bad in (2, -1):
        try:
            gp.div_correction(bad)
            raise AssertionError("group out of range was accepted")
        except M.MHDFlowsError:
            pass
    gp.close()


@case
def div_corrections_float32():
    _div_case(np.float32, F32_TOL)


@case
def div_corrections_float64():
    _div_case(np.float64, F64_TOL)


@case
def div_b_correction_emhd_and_hd_refusal():
    kw = dict(nx=16, T=np.float32, B_field=True, EMHD=True, dt=2e-4)
    op, gp = O.Problem(**kw), M.Problem(M.GPU(), **kw)
    g = op.grid
    rng = np.random.default_rng(9)
    f = [_band_limited(g, rng.standard_normal((16, 16, 16)).astype(np.float32)) for _ in range(3)]
    O.SetUpProblemIC(op, bx=f[0], by=f[1], bz=f[2])
    M.SetUpProblemIC(gp, bx=f[0], by=f[1], bz=f[2])
    FO.DivBCorrection(op)
    M.DivBCorrection(gp)
    assert O.rel_l2(gp.sol, _dealiased(op)) < F32_TOL
    O.stepforward(op)
    M.stepforward(gp)
    assert O.rel_l2(gp.sol, _dealiased(op)) < F32_TOL      # the (B.grad)A term read the refreshed stale b
    try:
        M.DivVCorrection(gp)
        raise AssertionError("DivVCorrection! on an EMHD problem was accepted")
    except M.MHDFlowsError:
        pass
    gp.close()
    hd = M.Problem(M.GPU(), nx=16)
    try:
        M.DivBCorrection(hd)
        raise AssertionError("DivBCorrection! on an HD problem was accepted")
    except M.MHDFlowsError:
        pass
    hd.close()


def _vp_case(B, T, tol, steps):
    op, gp = _vp_pair(M, O, B, T, dims=DIMS)
    g = op.grid
    N = np.zeros_like(op.sol)
    op.calcN(N, op.sol.copy(), 0.0, op.clock, op.vars, op.params, g)
    err = O.rel_l2(gp.calcN(), g.dealias(N.copy()))
    assert err < tol, err
    q = O.Problem(nx=g.nx, ny=g.ny, nz=g.nz, T=T, nu=2e-2, dt=2e-3, **(dict(eta=3e-2, B_field=True) if B else {}))
    q.sol[...] = op.sol
    N0 = np.zeros_like(op.sol)
    q.calcN(N0, q.sol.copy(), 0.0, q.clock, q.vars, q.params, g)
    assert O.rel_l2(g.dealias(N.copy()), g.dealias(N0.copy())) > 1e-2
    for _ in range(steps):
        O.stepforward(op)
    M.stepforward(gp, steps)
    assert O.rel_l2(gp.sol, _dealiased(op)) < tol
    gp.close()


@case
def volume_penalisation_hd_float32():
    _vp_case(False, np.float32, F32_TOL, 2)


@case
def volume_penalisation_mhd_float32():
    _vp_case(True, np.float32, F32_TOL, 1)


@case
def volume_penalisation_mhd_float64():
    _vp_case(True, np.float64, F64_TOL, 1)


@case
def volume_penalisation_time_integrator_and_refusals():
    op, gp = _vp_pair(M, O, True, np.float32, dims=(16, 16, 16))
    O.TimeIntegrator(op, 1e9, 1, usr_dt=1.5e-3)
    M.TimeIntegrator(gp, 1e9, 1, usr_dt=1.5e-3)
    g = op.grid
    assert gp.clock.step == op.clock.step == 2
    assert O.rel_l2(gp.sol, _dealiased(op)) < F32_TOL
    sol = gp.sol
    for base in (0, 3):
        div = g.kr * sol[base] + g.l * sol[base + 1] + g.m * sol[base + 2]
        assert np.linalg.norm(div.ravel()) / np.linalg.norm(sol[base:base + 3].ravel()) < 1e-5
    gp.close()
    try:
        M.Problem(M.GPU(), nx=16, B_field=True, EMHD=True, VP_method=True)
        raise AssertionError("VP_method + EMHD was accepted")
    except ValueError:
        pass
    plain = M.Problem(M.GPU(), nx=16)
    try:
        plain.set_vp_field("χ", np.zeros((16, 16, 16), np.float32))
        raise AssertionError("set_vp_field without VP_method was accepted")
    except M.MHDFlowsError:
        pass
    plain.close()


def _emhd_run(flag, stepper, T):
    os.environ["MHDF_EMHD2"] = flag
    p = M.Problem(M.GPU(), nx=16, ny=16, nz=16 if stepper == "LSRK54" else 32, T=T, stepper=stepper, B_field=True, EMHD=True, dt=1e-4)
    rng = np.random.default_rng(11)
    f = [rng.standard_normal(p._real_shape).astype(T) for _ in range(3)]
    M.SetUpProblemIC(p, bx=f[0], by=f[1], bz=f[2])
    M.stepforward(p)
    out = (p.sol, p.get_real("by", M.STALE), p.stale_stats()[0])
    p.close()
    os.environ.pop("MHDF_EMHD2", None)
    return out


@case
def second_emhd_kernel_form_is_bit_identical():
    for stepper, T in (("RK4", np.float32), ("LSRK54", np.float64)):
        a, b = _emhd_run("0", stepper, T), _emhd_run("1", stepper, T)
        assert np.linalg.norm(a[0]) > 0
        for x, y in zip(a, b):
            assert np.array_equal(x, y), stepper


@case
def hm89_stepper_emhd_both_precisions():
    """HM89TimeStepper (timestepper/HM89.jl): predictor, fixed-point loop with the stale-b quirk, resistive stages, CFL inputs."""
    _hm89_check(M, O, np.float64, F64_TOL, (16, 16, 16), steps=2)
    _hm89_check(M, O, np.float32, F32_TOL, (16, 16, 32), steps=1)


@case
def hdf5_dump_and_restart_through_the_library():
    kw = dict(nx=16, T=np.float32, nu=2e-2, eta=3e-2, dt=2e-3, B_field=True)
    gp = M.Problem(M.GPU(), **kw)
    ic = O.taylor_green_ic(O.Grid(16, T=np.float32))
    M.SetUpProblemIC(gp, ux=ic[0], uy=ic[1], uz=ic[2], bx=ic[3], by=ic[4], bz=ic[5])
    with tempfile.TemporaryDirectory() as d:
        M.TimeIntegrator(gp, 1e9, 1, usr_dt=2e-3, save=True, save_loc=d + "/", filename="run", dump_dt=2e-3)
        files = sorted(os.listdir(d))
        assert files[0] == "run_t_0000.h5" and len(files) >= 2, files
        last = os.path.join(d, files[-1])
        dump = M.readMHDFlows(last)
        assert set(dump) == {"i_velocity", "j_velocity", "k_velocity", "i_mag_field", "j_mag_field", "k_mag_field", "time"}
        assert dump["i_velocity"].dtype == np.float32 and dump["time"].dtype == np.float32
        tup = M.readMHDFlows(last, as_tuple=True)
        assert len(tup) == 7 and np.array_equal(tup[3], dump["i_mag_field"])
        rp = M.Problem(M.GPU(), **kw)
        M.Restart(rp, last)
        assert abs(rp.clock.t - float(dump["time"])) < 1e-9 and rp.clock.step == 0
        assert O.rel_l2(rp.get_real("ux", M.FRESH), dump["i_velocity"]) < 1e-6
        assert O.rel_l2(rp.get_real("bz", M.FRESH), dump["k_mag_field"]) < 1e-6
        rp.close()
    gp.close()


# ---- regression guard for the hardware-verified paths (tiny versions of tests/test_gpu_parity.py) -------------------
@case
def regress_calcN_hd_emhd_noncubic():
    for kind, dims in (("hd", (16, 16, 16)), ("emhd", (16, 16, 16)), ("mhd", (32, 16, 16))):
        op, gp = _pair(M, O, kind, dims, np.float32)
        N = np.zeros_like(op.sol)
        op.calcN(N, op.sol, 0.0, op.clock, op.vars, op.params, op.grid)
        op.grid.dealias(N)
        got = gp.calcN()
        for i in range(op.Nl):
            assert O.rel_l2(got[i], N[i]) < F32_TOL, (kind, i)
        gp.close()


@case
def regress_per_step_parity_lsrk54_f64_and_emhd():
    for kind, stepper, T, tol in (("hd", "LSRK54", np.float64, F64_TOL), ("emhd", "RK4", np.float32, F32_TOL), ("mhd", "LSRK54", np.float32, F32_TOL)):
        op, gp = _pair(M, O, kind, (16, 16, 16), T, stepper=stepper)
        for s in range(2):
            O.stepforward(op)
            M.stepforward(gp)
            assert O.rel_l2(gp.sol, _dealiased(op)) < tol, (kind, s)
        assert abs(gp.clock.t - op.clock.t) < 1e-6 and gp.clock.step == op.clock.step
        gp.close()


@case
def regress_cfl_path_stale_vars_spectrum_helicity():
    op, gp = _pair(M, O, "mhd", (16, 16, 16), np.float32, turb=False)
    O.TimeIntegrator(op, 1e9, 1, CFL_Coef=0.25)
    M.TimeIntegrator(gp, 1e9, 1, CFL_Coef=0.25)
    assert gp.clock.step == op.clock.step == 2
    assert abs(gp.clock.dt - op.clock.dt) / op.clock.dt < 1e-5
    assert O.rel_l2(gp.sol, _dealiased(op)) < 5e-5
    assert O.rel_l2(gp.vars.ux, op.vars.ux) < 1e-5 and O.rel_l2(gp.vars.bz, op.vars.bz) < 1e-5
    Pk, kr = M.spectralline(gp, "bx")
    Pk_ref, kr_ref = O.spectralline(op.grid.irfft(_dealiased(op)[3]), op.grid)
    assert len(Pk) == len(Pk_ref) and np.abs(Pk - Pk_ref).max() / Pk_ref.max() < 1e-4
    gp.close()
    # helicities / energies on a broadband field, reference evaluated in Float64 from the oracle's state
    op, gp = _pair(M, O, "mhd", (16, 16, 16), np.float32, turb=True)
    O.stepforward(op)
    M.stepforward(gp)
    g64 = O.Grid(16, T=np.float64)
    sol = _dealiased(op).astype(np.complex128)
    u = [g64.irfft(sol[i].copy()) for i in range(3)]
    b = [g64.irfft(sol[3 + i].copy()) for i in range(3)]
    hk_ref, hm_ref = float(np.sum(O.h_k(*u, g64))), float(np.sum(O.h_m(*b, g64)))
    hm_scale = float(np.sum(np.abs(O.h_m(*b, g64))))
    hk_scale = float(np.sum(np.abs(O.h_k(*u, g64))))      # the kinetic helicity of this field is a cancelling sum (1e-3 of its scale)
    hk, hm, hc = gp.helicity()
    dv = g64.dx * g64.dy * g64.dz
    hc_ref = float(np.sum(u[0] * b[0] + u[1] * b[1] + u[2] * b[2])) * dv
    assert abs(hc - hc_ref) < 1e-5 * abs(hc_ref) and abs(hk - hk_ref) < 1e-5 * hk_scale, (hk, hk_ref, hk_scale, hc, hc_ref)
    assert abs(hm - hm_ref) < 1e-5 * hm_scale, (hm, hm_ref, hm_scale)
    ke, me = gp.energy(M.FRESH)
    assert abs(ke - float(sum(np.sum(x ** 2) for x in u)) * dv) < 1e-5 * ke and abs(me - float(sum(np.sum(x ** 2) for x in b)) * dv) < 1e-5 * me
    gp.close()


@case
def regress_errors_and_n97_forcing():
    try:
        M.Problem(M.GPU(), nx=48)
        raise AssertionError("nx = 48 was accepted")
    except M.MHDFlowsError:
        pass
    p = M.Problem(M.GPU(), nx=16)
    try:
        p.get_spectral(5)
        raise AssertionError("field 5 of an HD problem was accepted")
    except M.MHDFlowsError:
        pass
    p.set_real(0, np.full((16, 16, 16), np.nan, dtype=np.float32))
    p.clock.dt = 1e-3
    try:
        M.stepforward(p)
        raise AssertionError("NaN went unnoticed")
    except M.MHDFlowsError as e:
        assert e.code == -4
    p.close()
    kw = dict(nx=16, T=np.float32, nu=2e-2, eta=3e-2, dt=4e-3, B_field=True)
    g0 = O.Grid(16, T=np.float32)
    X, Y, Z = (g0.x.astype(np.float64).reshape(1, 1, -1), g0.y.astype(np.float64).reshape(1, -1, 1), g0.z.astype(np.float64).reshape(-1, 1, 1))
    fxh = g0.rfft((0.5 * np.sin(2 * X) * np.cos(2 * Y) * np.cos(2 * Z)).astype(np.float32))
    fyh = g0.rfft((-0.5 * np.cos(2 * X) * np.sin(2 * Y) * np.cos(2 * Z)).astype(np.float32))

    def calcF(N, sol, t, clock, vars, params, grid):
        N[params.ux_ind] += fxh
        N[params.uy_ind] += fyh

    op = O.Problem(calcF=calcF, **kw)
    uv, fn = M.GetN97vars_And_function(M.GPU(), 16, 16, 16)
    gp = M.Problem(M.GPU(), calcF=fn, usr_vars=uv, **kw)
    u, b = O.random_phase_ic(op.grid, 11), O.random_phase_ic(op.grid, 12)
    O.SetUpProblemIC(op, *u, bx=b[0], by=b[1], bz=b[2])
    M.SetUpProblemIC(gp, ux=u[0], uy=u[1], uz=u[2], bx=b[0], by=b[1], bz=b[2])
    M.SetUpN97(gp, F0=0.5, kf=2)
    for _ in range(2):
        O.stepforward(op)
    M.stepforward(gp, 2)
    assert O.rel_l2(gp.sol, _dealiased(op)) < F32_TOL
    gp.close()


if __name__ == "__main__":
    only = sys.argv[1:]
    failed = 0
    for fn in CASES:
        if only and not any(o in fn.__name__ for o in only):
            continue
        t0 = time.time()
        try:
            fn()
            print(f"PASS {fn.__name__} ({time.time() - t0:.1f} s)", flush=True)
        except Exception:
            failed += 1
            print(f"FAIL {fn.__name__} ({time.time() - t0:.1f} s)\n{traceback.format_exc()}", flush=True)
    print(f"emu-lib cases done: {failed} failure(s)", flush=True)
    sys.exit(1 if failed else 0)
