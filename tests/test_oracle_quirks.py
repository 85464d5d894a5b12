"""The reference's quirks that the oracle (and therefore the CUDA path) must reproduce -- SURVEY App. A.4-A.7."""
import numpy as np

from oracle import fused_model as FM
from oracle import mhdflows_oracle as O


def _mhd(n=16, T=np.float64, **kw):
    p = O.Problem(nx=n, T=T, dt=4e-3, nu=0.02, eta=0.03, B_field=True, **kw)
    u, b = O.random_phase_ic(p.grid, 1), O.random_phase_ic(p.grid, 2)
    O.SetUpProblemIC(p, *u, bx=b[0], by=b[1], bz=b[2])
    return p


def test_stale_vars_hold_the_last_stage_input_not_sol():
    """A.5: after a step vars.u* = c2r of the 4th RK4 stage input; getCFL!/ProbDiagnostic read those."""
    p = _mhd()
    O.stepforward(p)
    fresh = p.grid.irfft(p.grid.dealias(p.sol[0].copy()))
    assert O.rel_l2(p.vars.ux, fresh) > 1e-10
    assert O.rel_l2(p.vars.ux, p.grid.irfft(p.timestepper.sol1[0].copy())) < 1e-14


def test_dealias_is_applied_in_place_to_the_stage_input_only():
    """A.2: calcN! masks its input in place; N and the post-step sol keep aliased-band garbage."""
    p = _mhd()
    p.sol += 1e-3            # pollute every mode, including the aliased band
    N = np.zeros_like(p.sol)
    p.calcN(N, p.sol, 0.0, p.clock, p.vars, p.params, p.grid)
    msk = p.grid.retained_mask()
    assert np.all(p.sol[:, ~msk] == 0)          # input masked in place
    assert np.abs(N[:, ~msk]).max() > 0         # output not masked


def test_truncation_is_asymmetric():
    """A.4: k = -N/3 is kept, +N/3 is dropped (N = 24: -8 kept, +8 dropped)."""
    g = O.Grid(24, T=np.float64)
    msk = g.retained_mask()
    l = np.rint(g.l.ravel() / (2 * np.pi / g.Ly)).astype(int)
    kept = sorted(l[msk[0, :, 0]])
    assert kept[0] == -8 and kept[-1] == 7


def test_hd_forcing_is_clobbered_but_mhd_forcing_acts():
    """A.7 / a5: HDcalcN! adds the forcing before the advection zeroes N."""
    def calcF(N, sol, t, clock, vars, params, grid):
        N[params.ux_ind] += 1.0

    for B, acts in ((False, False), (True, True)):
        kw = dict(nx=16, T=np.float64, dt=1e-3, nu=0.01)
        if B:
            kw.update(eta=0.01, B_field=True)
        a, b = O.Problem(calcF=calcF, **kw), O.Problem(**kw)
        u = O.random_phase_ic(a.grid, 3)
        for q in (a, b):
            O.SetUpProblemIC(q, *u, **(dict(bx=u[0], by=u[1], bz=u[2]) if B else {}))
        Na, Nb = np.zeros_like(a.sol), np.zeros_like(b.sol)
        a.calcN(Na, a.sol, 0.0, a.clock, a.vars, a.params, a.grid)
        b.calcN(Nb, b.sol, 0.0, b.clock, b.vars, b.params, b.grid)
        assert (np.abs(Na - Nb).max() > 0.5) == acts


def test_hyperviscosity_adds_on_top_of_viscosity():
    """A.7: n_nu > 1 gives -nu k^2 u - nu k^(2 n_nu) u, not a replacement."""
    kw = dict(nx=16, T=np.float64, dt=1e-3)
    a, b = O.Problem(nu=0.01, n_nu=2, **kw), O.Problem(nu=0.01, n_nu=0, **kw)
    u = O.random_phase_ic(a.grid, 4)
    for q in (a, b):
        O.SetUpProblemIC(q, *u)
    Na, Nb = np.zeros_like(a.sol), np.zeros_like(b.sol)
    a.calcN(Na, a.sol, 0.0, a.clock, a.vars, a.params, a.grid)
    b.calcN(Nb, b.sol, 0.0, b.clock, b.vars, b.params, b.grid)
    g = a.grid
    k2 = g.Krsq.astype(np.float64)
    us = FM.sym_kr0(g.dealias(a.sol.copy())[0], g)
    msk = g.retained_mask()
    assert O.rel_l2((Na[0] - Nb[0])[msk], (-(k2 ** 2) * 0.01 * us)[msk]) < 1e-12


def test_emhd_uses_the_stale_b():
    """A.6 (i): the (B.grad)A term reads vars.b* of the PREVIOUS calcN! call; a fresh-b evaluation deviates."""
    p = O.Problem(nx=16, T=np.float64, dt=5e-4, B_field=True, EMHD=True)
    b = O.random_phase_ic(p.grid, 5)
    O.SetUpProblemIC(p, bx=b[0], by=b[1], bz=b[2])
    f = FM.FusedProblem(p)
    for _ in range(2):
        O.stepforward(p)
        f.step()
    assert O.rel_l2(f.sol, p.grid.dealias(p.sol.copy())) < 1e-13
    # fresh-b variant: refresh the stale field before every evaluation
    q = O.Problem(nx=16, T=np.float64, dt=5e-4, B_field=True, EMHD=True)
    O.SetUpProblemIC(q, bx=b[0], by=b[1], bz=b[2])
    h = FM.FusedProblem(q)
    orig = h.rhs

    def fresh_rhs(s):
        g = h.grid
        sm = g.dealias(s.copy())
        h.b_stale = [g.irfft(sm[i].copy()) for i in range(3)]
        return orig(s)

    h.rhs = fresh_rhs
    for _ in range(2):
        h.step()
    assert O.rel_l2(h.sol, p.grid.dealias(p.sol.copy())) > 1e-8


def test_time_integrator_resets_step_counter_and_uses_t_diff():
    p = _mhd(T=np.float32)
    p.clock.step = 17
    O.TimeIntegrator(p, 1e9, 0, CFL_Coef=0.25)
    assert p.clock.step == 1                                   # reset to 0, then N0 + 1 = 1 step
    dl = 2 * np.pi / 16
    assert p.clock.dt <= 0.25 * dl * dl / 0.03 * (1 + 1e-6)    # t_diff = CFL_Coef dl^2 / max(nu, eta)
