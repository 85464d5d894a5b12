"""Known-answer physics tests that pin the oracle's reading of the reference's signs and terms
(SURVEY.md section 4.2).  The reference has no tests of its own; these replace them."""
import math

import numpy as np
import pytest

from oracle import mhdflows_oracle as O


def _grid_xyz(g):
    X = g.x.astype(np.float64).reshape(1, 1, -1)
    Y = g.y.astype(np.float64).reshape(1, -1, 1)
    Z = g.z.astype(np.float64).reshape(-1, 1, 1)
    return X, Y, Z


def test_alias_ranges_match_julia_float64_expressions():
    # SURVEY App. A.2 table
    assert O.aliased_range(32) == (11, 22)
    assert O.aliased_range(96) == (33, 64)
    assert O.aliased_range(256) == (86, 171)
    assert O.aliased_range(512) == (171, 342)
    assert O.aliased_range(1024) == (342, 683)


def test_fft_count_per_rhs():
    """36 / 24 / 51 full 3D FFTs per RHS evaluation."""
    for kw, expect in ((dict(B_field=True), 36), (dict(), 24), (dict(B_field=True, EMHD=True), 51)):
        p = O.Problem(nx=16, T=np.float64, dt=1e-3, **kw)
        cnt = [0]
        rf, irf = p.grid.rfft, p.grid.irfft
        p.grid.rfft = lambda f: (cnt.__setitem__(0, cnt[0] + 1), rf(f))[1]
        p.grid.irfft = lambda f: (cnt.__setitem__(0, cnt[0] + 1), irf(f))[1]
        N = np.zeros_like(p.sol)
        p.calcN(N, p.sol, 0.0, p.clock, p.vars, p.params, p.grid)
        assert cnt[0] == expect


def test_abc_flow_decays_exactly():
    """Beltrami (ABC) flow: u x omega = 0, so u(t) = u(0) exp(-nu t): pins projection+diffusion+RK4."""
    nu, dt, nsteps = 0.05, 0.01, 100
    p = O.Problem(nx=32, T=np.float64, nu=nu, dt=dt)
    X, Y, Z = _grid_xyz(p.grid)
    A, B, C = 1.0, 0.7, 0.4
    ux = A * np.sin(Z) + C * np.cos(Y) + 0 * X
    uy = B * np.sin(X) + A * np.cos(Z) + 0 * Y
    uz = C * np.sin(Y) + B * np.cos(X) + 0 * Z
    O.SetUpProblemIC(p, ux=ux, uy=uy, uz=uz)
    sol0 = p.sol.copy()
    for _ in range(nsteps):
        O.stepforward(p)
    # RK4 on y' = -nu y has its own (tiny) truncation error: compare with the RK4 amplification factor
    z = -nu * dt
    amp = (1 + z + z * z / 2 + z ** 3 / 6 + z ** 4 / 24) ** nsteps
    assert O.rel_l2(p.sol, sol0 * amp) < 1e-13
    assert abs(amp - math.exp(-nu * dt * nsteps)) < 1e-12


def test_shear_alfven_wave():
    """b = B0 x^ + b_y, u_y(0) = eps sin(kx): exact nonlinear solution for nu = eta."""
    B0, k, nu, eps, dt, nsteps = 1.5, 3, 0.02, 1e-3, 0.005, 100
    p = O.Problem(nx=32, T=np.float64, nu=nu, eta=nu, dt=dt, B_field=True)
    X, Y, Z = _grid_xyz(p.grid)
    zero = 0 * X + 0 * Y + 0 * Z
    O.SetUpProblemIC(p, ux=zero, uy=eps * np.sin(k * X) + zero, uz=zero, bx=B0 + zero, by=zero, bz=zero)
    for _ in range(nsteps):
        O.stepforward(p)
    t = dt * nsteps
    uy = eps * math.cos(k * B0 * t) * math.exp(-nu * k * k * t) * np.sin(k * X) + zero
    by = eps * math.sin(k * B0 * t) * math.exp(-nu * k * k * t) * np.cos(k * X) + zero
    g = p.grid
    assert O.rel_l2(g.irfft(p.sol[1].copy()), uy) < 1e-7
    assert O.rel_l2(g.irfft(p.sol[4].copy()), by) < 1e-7


def test_emhd_whistler_wave():
    """B = B0 z^ + eps (cos kz, sin kz, 0) rotates with phase +k^2 B0 t (d_i = 1)."""
    B0, k, eps, dt, nsteps = 1.0, 2, 1e-3, 0.00125, 400
    p = O.Problem(nx=16, T=np.float64, dt=dt, B_field=True, EMHD=True)
    X, Y, Z = _grid_xyz(p.grid)
    zero = 0 * X + 0 * Y + 0 * Z
    O.SetUpProblemIC(p, bx=eps * np.cos(k * Z) + zero, by=eps * np.sin(k * Z) + zero, bz=B0 + zero)
    for _ in range(nsteps):
        O.stepforward(p)
    phi = k * k * B0 * dt * nsteps
    g = p.grid
    bx = eps * np.cos(k * Z + phi) + zero
    by = eps * np.sin(k * Z + phi) + zero
    assert O.rel_l2(g.irfft(p.sol[0].copy()), bx) < 1e-8
    assert O.rel_l2(g.irfft(p.sol[1].copy()), by) < 1e-8
    # the opposite rotation sense must fail at O(1)
    assert O.rel_l2(g.irfft(p.sol[0].copy()), eps * np.cos(k * Z - phi) + zero) > 0.5


def test_ideal_mhd_invariants_and_solenoidality():
    p = O.Problem(nx=24, T=np.float64, dt=2e-3, B_field=True)
    g = p.grid
    ic = O.taylor_green_ic(g)
    O.SetUpProblemIC(p, *ic[:3], bx=ic[3], by=ic[4], bz=ic[5])

    def invariants():
        u = [g.irfft(p.sol[i].copy()).astype(np.float64) for i in range(3)]
        b = [g.irfft(p.sol[i].copy()).astype(np.float64) for i in range(3, 6)]
        E = sum(np.sum(a * a) for a in u + b)
        Hc = sum(np.sum(a * c) for a, c in zip(u, b))
        return E, Hc

    E0, H0 = invariants()
    for _ in range(20):
        O.stepforward(p)
    E1, H1 = invariants()
    assert abs(E1 - E0) / E0 < 1e-8
    assert abs(H1 - H0) < 1e-8 * E0
    msk = g.retained_mask()
    for base in (0, 3):
        div = g.kr * p.sol[base] + g.l * p.sol[base + 1] + g.m * p.sol[base + 2]
        nrm = np.sqrt(sum(np.sum(np.abs(p.sol[base + i][msk]) ** 2) for i in range(3)))
        assert np.sqrt(np.sum(np.abs(div[msk]) ** 2)) / nrm < 1e-12


def test_lsrk54_tableau_order_conditions():
    from fractions import Fraction as F
    A, B, C = O.LSRK54_A, O.LSRK54_B, O.LSRK54_C
    # Butcher form of a 2N-storage scheme: a_{i,j} = sum_{m=j}^{i-1} B_m prod_{l=j+1}^{m} A_l
    s = 5
    a = [[F(0)] * s for _ in range(s)]
    for i in range(s):
        for j in range(i):
            tot = F(0)
            for m in range(j, i):
                prod = F(1)
                for l in range(j + 1, m + 1):
                    prod *= A[l]
                tot += B[m] * prod
            a[i][j] = tot
    b = [sum(B[m] * math.prod([A[l] for l in range(j + 1, m + 1)] or [F(1)]) for m in range(j, s)) for j in range(s)]
    c = [sum(a[i]) for i in range(s)]
    assert all(abs(float(c[i] - C[i])) < 1e-12 for i in range(s))
    f = lambda x: abs(float(x))
    assert f(sum(b) - 1) < 1e-12
    assert f(sum(bi * ci for bi, ci in zip(b, c)) - F(1, 2)) < 1e-12
    assert f(sum(bi * ci ** 2 for bi, ci in zip(b, c)) - F(1, 3)) < 1e-12
    assert f(sum(b[i] * a[i][j] * c[j] for i in range(s) for j in range(s)) - F(1, 6)) < 1e-12
    assert f(sum(bi * ci ** 3 for bi, ci in zip(b, c)) - F(1, 4)) < 1e-12
    assert f(sum(b[i] * c[i] * a[i][j] * c[j] for i in range(s) for j in range(s)) - F(1, 8)) < 1e-12
    assert f(sum(b[i] * a[i][j] * c[j] ** 2 for i in range(s) for j in range(s)) - F(1, 12)) < 1e-12
    assert f(sum(b[i] * a[i][j] * a[j][k] * c[k] for i in range(s) for j in range(s) for k in range(s)) - F(1, 24)) < 1e-12


def test_lsrk54_matches_rk4_to_fourth_order():
    res = {}
    for stepper in ("RK4", "LSRK54"):
        p = O.Problem(nx=16, T=np.float64, dt=5e-3, nu=0.01, eta=0.01, B_field=True, stepper=stepper)
        ic = O.taylor_green_ic(p.grid)
        O.SetUpProblemIC(p, *ic[:3], bx=ic[3], by=ic[4], bz=ic[5])
        for _ in range(10):
            O.stepforward(p)
        p.grid.dealias(p.sol)
        res[stepper] = p.sol.copy()
    assert O.rel_l2(res["LSRK54"], res["RK4"]) < 1e-9


def test_time_integrator_runs_n0_plus_one_steps_and_cfl():
    p = O.Problem(nx=16, T=np.float32, nu=0.01, eta=0.01, B_field=True)
    ic = O.taylor_green_ic(p.grid)
    O.SetUpProblemIC(p, *ic[:3], bx=ic[3], by=ic[4], bz=ic[5])
    O.TimeIntegrator(p, 1e9, 3, CFL_Coef=0.25)
    assert p.clock.step == 4  # integrator.jl:104 off-by-one quirk
    vmax = 2.0  # max |bz| of the TG field
    assert abs(p.clock.dt - 0.25 * (2 * math.pi / 16) / vmax) / p.clock.dt < 0.2
    KE, ME = O.ProbDiagnostic(p)
    assert KE > 0 and ME > 0 and KE == O.round_sig(KE, 3)


def test_spectralline_parseval_half_spectrum():
    g = O.Grid(16, T=np.float64)
    rng = np.random.default_rng(0)
    A = rng.standard_normal((16, 16, 16))
    Pk, kr = O.spectralline(A, g)
    assert len(Pk) == int(np.rint(math.sqrt(3) * 8 + 1))
    assert abs(Pk.sum() - np.sum(np.abs(g.rfft(A)) ** 2)) / Pk.sum() < 1e-12


def test_divfree_spectra_map_is_solenoidal_and_real():
    g = O.Grid(24, T=np.float64)
    fx, fy, fz = O.random_phase_ic(g, seed=1234)
    div = g.kr * g.rfft(fx) + g.l * g.rfft(fy) + g.m * g.rfft(fz)
    nrm = np.sqrt(np.sum(np.abs(g.rfft(fx)) ** 2 + np.abs(g.rfft(fy)) ** 2 + np.abs(g.rfft(fz)) ** 2))
    # solenoidal up to the Hermitian symmetrisation the c2r applies on the kr=0 plane (zeroed anyway)
    assert np.sqrt(np.sum(np.abs(div) ** 2)) / nrm < 1e-10
    assert np.isfinite(fx).all() and fx.std() > 0


def test_hm89_restatement_properties():
    """timestepper/HM89.jl, as written: the field leaves a step solenoidal (DivFreeCorrection!), the fixed point converges below
    5e-4, and -- because RK3linearterm! reads the arrays the loop left behind (calcF! = nothingfunction never clears them) -- the
    resistive stages re-apply the Hall term: with eta = 0 a step advances B by (1 + 15/16 - 8/15 * 153/128) dt N = 1.3 dt N(B_half)."""
    kw = dict(nx=16, dt=1e-3, eta=0.0, B_field=True, EMHD=True, T=np.float64)
    p = O.Problem(stepper="HM89", **kw)
    g = p.grid
    b = [g.irfft(g.dealias(g.rfft(x))) for x in O.random_phase_ic(g, 3)]
    O.SetUpProblemIC(p, bx=b[0], by=b[1], bz=b[2])
    B0 = p.sol.copy()
    O.stepforward(p)
    ts = p.timestepper
    assert 2 <= ts.iters < 20 and ts.eps <= 5e-4
    assert p.clock.step == 1 and p.clock.t == 1e-3
    m = g.retained_mask()
    div = g.kr * p.sol[0] + g.l * p.sol[1] + g.m * p.sol[2]
    assert np.abs(div[m]).max() < 1e-13 * np.abs(p.sol[:, m]).max()
    q = O.Problem(stepper="RK4", **kw)            # scratch vars for one more evaluation; stale b = the last (tiny) difference ~ 0
    N = np.zeros_like(B0)
    O.EMHDcalcN(N, (B0 + ts.B1) * 0.5, 0, q.clock, q.vars, q.params, g)
    pred = B0 + (1 + 15 / 16 - 8 / 15 * 153 / 128) * 1e-3 * N
    O.DivFreeCorrection(pred, q.vars, q.params, g)
    assert O.rel_l2((p.sol - B0)[:, m], (pred - B0)[:, m]) < 2e-2
    with pytest.raises(ValueError):
        O.Problem(nx=16, B_field=True, stepper="HM89")          # Problems.jl:124: only with EFlag


def test_correlation_and_structure_function_restatements():
    """TurbStatTool.jl:67-120: CF is the periodic autocorrelation sum_x V(x) V(x + r) with zero lag at the fftshift centre; SFC and
    SF_2 1D follow it literally (mean of V, shells measured from element N/2 in 1-based counting)."""
    rng = np.random.default_rng(5)
    V = rng.standard_normal((8, 8, 8))
    cf = O.CF(V)
    for r in ((0, 0, 0), (1, 0, 0), (0, 3, 0), (2, 5, 7)):
        direct = float(np.sum(V * np.roll(V, shift=tuple(-x for x in r), axis=(0, 1, 2))))
        assert abs(cf[(4 + r[0]) % 8, (4 + r[1]) % 8, (4 + r[2]) % 8] - direct) < 1e-10 * abs(cf[4, 4, 4])
    assert np.allclose(O.SFC(V), 2 * (V.mean() - cf))
    sf = O.SF2_1D(V, V, V)
    R = math.ceil(math.sqrt(3 * 4 ** 2))
    assert sf.shape == (2 * R,) and np.isfinite(sf[:6]).all() and np.isnan(sf[R + 1:]).all()
    # shell 1 of the literal loop: round(|r|) = 1 (distances 1 and sqrt 2: 18 elements) around element (N/2, N/2, N/2) in 1-based
    # counting = 0-based (3, 3, 3) -- one off the zero lag at (4, 4, 4)
    s3 = 3 * O.SFC(V).astype(np.float64)
    ax = np.arange(8) - 3
    d = np.rint(np.sqrt(ax.reshape(-1, 1, 1) ** 2 + ax.reshape(1, -1, 1) ** 2 + ax.reshape(1, 1, -1) ** 2))
    assert (d == 1).sum() == 18
    assert abs(sf[0] - s3[d == 1].mean()) < 1e-9 * abs(s3[d == 1].mean())
