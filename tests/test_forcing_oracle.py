"""CPU tests of oracle/forcing_oracle.py: the Philox generator against the Random123 known-answer vectors, the A99
forcing restatements (solenoidal basis, quirks of the two reference implementations) and the divergence corrections."""
import numpy as np
import pytest

from oracle import forcing_oracle as FO
from oracle import mhdflows_oracle as O


def test_philox4x32_known_answers():
    # Random123 kat_vectors: philox4x32-10
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, out in kat:
        got = FO.philox4x32(ctr, key)
        assert tuple(int(x) for x in got) == out
    # vectorised == scalar
    c0 = np.arange(5, dtype=np.uint64)
    v = FO.philox4x32((c0, 7, 8, 9), (1, 2))
    for i in range(5):
        s = FO.philox4x32((i, 7, 8, 9), (1, 2))
        assert [int(x[i]) for x in v] == [int(x) for x in s]


@pytest.mark.parametrize("T", [np.float32, np.float64])
def test_uniform_stream_properties(T):
    g = O.Grid(16, 16, 32, T=T)
    rng = FO.PhiloxField(1234, g)
    r = rng.uniforms(0)
    assert all(x.dtype == T and x.shape == (32, 16, 9) for x in r)
    assert all(0 <= x.min() and x.max() < 1 for x in r)
    allr = np.concatenate([x.ravel() for x in r])
    assert abs(allr.mean() - 0.5) < 0.01 and abs(allr.var() - 1 / 12) < 0.005
    r2 = rng.uniforms(1)
    assert not np.array_equal(r[0], r2[0])                 # a new call draws new numbers
    assert np.array_equal(rng.uniforms(0)[2], r[2])        # counter based: reproducible


def _forced_problem(T, variant, seed=99):
    if variant == "host":
        p = O.Problem(nx=16, ny=16, nz=32, T=T, nu=1e-2, eta=1e-2, dt=1e-3, B_field=True, calcF=FO.A99ForceDriving)
        p.vars.usr_vars = FO.A99Vars(p.grid)
        FO.SetUpFk(p, kf=2, P=1, sigma2=1)
    else:
        p = O.Problem(nx=16, ny=16, nz=32, T=T, nu=1e-2, eta=1e-2, dt=1e-3, B_field=True, calcF=FO.A99ForceDriving_GPU)
        p.vars.usr_vars = FO.A99GPUVars(p.grid)
        FO.SetUpFk_GPU(p, kf=2.0, P=1.0, sigma=1.0, b=1.0)
    p.vars.usr_vars.rng = FO.PhiloxField(seed, p.grid)
    return p


@pytest.mark.parametrize("variant", ["host", "gpu"])
def test_a99_forcing_structure(variant):
    p = _forced_problem(np.float64, variant)
    g = p.grid
    N = np.zeros_like(p.sol)
    p.params.calcF(N, p.sol, 0.0, p.clock, p.vars, p.params, g)
    assert p.vars.usr_vars.calls == 1
    assert np.all(N[3:] == 0)                              # only the velocity is driven
    f = N[:3]
    assert np.abs(f).max() > 0
    div = g.kr * f[0] + g.l * f[1] + g.m * f[2]
    if variant == "host":
        assert np.all(f[:, :, :, 0] == 0)                  # Fk[1,:,:] = 0: no forcing on the kr = 0 plane
        # e1 acts on the first z plane only (copyto! of a (nkr, nl, 1) table); both basis vectors are orthogonal to k
        assert np.abs(div).max() < 1e-12 * np.abs(f).max() * np.abs(g.m).max()
        e1_part = np.abs(p.vars.usr_vars.e1x).sum(axis=(1, 2))
        assert e1_part[0] > 0 and np.all(e1_part[1:] == 0)
        # complex g_i: the second amplitude is a complex number as well
        assert np.abs(p.vars.usr_vars.gi.imag).max() > 1e-3
    else:
        inner = div[:, :, 1:-1]
        assert np.abs(inner).max() < 1e-12 * np.abs(f).max() * np.abs(g.m).max()
        assert np.all(f[:, :, :, 0].imag == 0) and np.all(f[:, :, :, -1].imag == 0)
    # a second evaluation draws different phases
    N2 = np.zeros_like(p.sol)
    p.params.calcF(N2, p.sol, 0.0, p.clock, p.vars, p.params, g)
    assert O.rel_l2(N2[:3], f) > 0.5


def test_a99_forced_run_injects_energy():
    p = _forced_problem(np.float32, "host")
    q = O.Problem(nx=16, ny=16, nz=32, T=np.float32, nu=1e-2, eta=1e-2, dt=1e-3, B_field=True)
    ic = O.taylor_green_ic(p.grid)
    for pr in (p, q):
        O.SetUpProblemIC(pr, *ic[:3], bx=ic[3], by=ic[4], bz=ic[5])
    for _ in range(5):
        O.stepforward(p)
        O.stepforward(q)
    assert p.vars.usr_vars.calls == 20                     # one forcing call per RK4 stage
    assert O.rel_l2(p.grid.dealias(p.sol.copy()), q.grid.dealias(q.sol.copy())) > 1e-4
    assert np.all(np.isfinite(p.sol))


@pytest.mark.parametrize("T,tol", [(np.float32, 1e-6), (np.float64, 1e-14)])
def test_div_corrections(T, tol):
    p = O.Problem(nx=16, T=T, nu=1e-2, eta=1e-2, dt=1e-3, B_field=True)
    g = p.grid
    rng = np.random.default_rng(5)
    fields = [rng.standard_normal((16, 16, 16)).astype(T) for _ in range(6)]     # not solenoidal
    O.SetUpProblemIC(p, *fields[:3], bx=fields[3], by=fields[4], bz=fields[5])
    before = p.sol.copy()
    FO.DivBCorrection(p)
    assert np.array_equal(p.sol[:3], before[:3])
    div = lambda s: g.kr * s[0] + g.l * s[1] + g.m * s[2]
    assert np.abs(div(before[3:])).max() > 1
    assert np.abs(div(p.sol[3:])).max() < tol * np.abs(before).max() * 16
    assert O.rel_l2(p.vars.bx, g.irfft(p.sol[3].copy())) < 10 * tol        # vars refreshed from the corrected sol
    # projection: applying it twice changes nothing; the solenoidal part is untouched
    once = p.sol.copy()
    FO.DivBCorrection(p)
    assert O.rel_l2(p.sol[3:], once[3:]) < 10 * tol
    FO.DivVCorrection(p)
    assert np.abs(div(p.sol[:3])).max() < tol * np.abs(before).max() * 16
    k2 = g.Krsq
    sol_part = before[:3] - np.stack([[g.kr, g.l, g.m][i] * div(before[:3]) * g.invKrsq for i in range(3)])
    assert O.rel_l2(p.sol[:3], sol_part.astype(g.CT)) < 10 * tol
    assert k2.shape == (16, 16, 9)


@pytest.mark.parametrize("B", [False, True])
def test_volume_penalisation_terms(B):
    """VP_method (VPSolver.jl:21-59 via HDSolver.jl:77-79 / MHDSolver.jl:86-88,161-163): chi = 0 leaves the equations
    untouched; a solid region damps the velocity towards the wall velocity; the added term equals
    -P[F(chi/eta (u - U0))] with P the solenoidal projector and eta = dt 13/7."""
    kw = dict(nx=16, T=np.float64, nu=1e-2, eta=1e-2, dt=1e-3, B_field=B)
    plain, zero, pen = O.Problem(**kw), O.Problem(VP_method=True, **kw), O.Problem(VP_method=True, **kw)
    g = plain.grid
    ic = O.taylor_green_ic(g)
    for p in (plain, zero, pen):
        O.SetUpProblemIC(p, *ic[:3], **(dict(bx=ic[3], by=ic[4], bz=ic[5]) if B else {}))
    chi = ((g.x.reshape(1, 1, -1) ** 2 + g.y.reshape(1, -1, 1) ** 2) < 1.0) * np.ones((16, 16, 16))
    pen.params.vp.chi[...] = chi
    pen.params.vp.U0x[...] = 0.25
    # the term itself, on one RHS evaluation
    N0, N1 = np.zeros_like(plain.sol), np.zeros_like(pen.sol)
    plain.calcN(N0, plain.sol.copy(), 0.0, plain.clock, plain.vars, plain.params, g)
    pen.calcN(N1, pen.sol.copy(), 0.0, pen.clock, pen.vars, pen.params, g)
    eta = 1e-3 * 13 / 7
    u = [g.irfft(g.dealias(plain.sol[i].copy())) for i in range(3)]
    W = [0.25, 0.0, 0.0]
    V = np.stack([g.rfft(chi / eta * (u[j] - W[j])) for j in range(3)])
    kV = (g.kr * V[0] + g.l * V[1] + g.m * V[2]) * g.invKrsq
    for a_, k in enumerate((g.kr, g.l, g.m)):
        assert O.rel_l2(N1[a_] - N0[a_], -(V[a_] - k * kV)) < 1e-12
    if B:
        assert O.rel_l2(N1[3:], N0[3:]) > 1e-3          # B0 = 0 walls: the induction equation is penalised as well
    for _ in range(5):
        for p in (plain, zero, pen):
            O.stepforward(p)
    assert O.rel_l2(zero.sol, plain.sol) == 0.0
    inside = lambda p: float(np.sum((chi * (p.vars.ux - 0.25)) ** 2))
    assert inside(pen) < 0.2 * inside(plain)
