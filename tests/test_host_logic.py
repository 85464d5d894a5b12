"""Host-side mirror of the reference API (mhdflows_jl_b200/problem.py): argument checking, grid formulas, the Diagnostic
recorder and the IC generator -- everything that does not need the device."""
import math

import numpy as np
import pytest

import mhdflows_jl_b200 as M
from mhdflows_jl_b200 import problem as P
from oracle import mhdflows_oracle as O


def test_problem_rejects_what_the_reference_rejects_before_touching_the_device():
    with pytest.raises(ValueError, match="Shear"):
        M.Problem(M.GPU(), nx=32, Shear=True)                       # pgen.jl:103-105
    with pytest.raises(ValueError, match="c"):
        M.Problem(M.GPU(), nx=32, Compressibility=True)             # pgen.jl:98-100
    with pytest.raises(ValueError, match="EMHD requires"):
        M.Problem(M.GPU(), nx=32, EMHD=True)
    with pytest.raises(ValueError, match="EMHD problems only"):
        M.Problem(M.GPU(), nx=32, stepper="HM89")                   # Problems.jl:124-126: HM89TimeStepper only with EFlag
    with pytest.raises(NotImplementedError, match="forcing"):
        M.Problem(M.GPU(), nx=32, B_field=True, EMHD=True, stepper="HM89", calcF=M.N97ForceDriving)
    with pytest.raises(ValueError):
        M.Problem(M.GPU(), nx=32, stepper="ETDRK4")
    with pytest.raises(M.MHDFlowsError):
        M.Problem(M.CPU(), nx=32)
    with pytest.raises(TypeError):
        M.Problem(M.GPU(), nx=32, bogus=1)


@pytest.mark.parametrize("dims", [(32, 32, 32), (64, 32, 16), (96, 48, 24)])
def test_grid_matches_oracle_grid(dims):
    nx, ny, nz = dims
    g = P._Grid(nx, ny, nz, 2 * math.pi, 3.0, 5.0, np.float32)
    o = O.Grid(nx, ny, nz, 2 * math.pi, 3.0, 5.0, np.float32)
    for a in ("x", "y", "z", "kr", "l", "m", "Krsq"):
        assert np.array_equal(getattr(g, a), getattr(o, a)), a
    assert np.array_equal(g.retained_mask(), o.retained_mask())
    for n in (16, 32, 96, 256, 512, 1024):
        assert g.aliased_range(n) == O.aliased_range(n)


class _FakeClock:
    def __init__(self):
        self.t, self.step, self.dt = 0.0, 0, 0.1


class _FakeProb:
    def __init__(self):
        self.clock = _FakeClock()


def test_diagnostic_recorder_semantics():
    """DiagnosticWrapper.jl:40-105: first value at construction, update every `freq` steps, auto-extend."""
    p = _FakeProb()
    d = M.Diagnostic(lambda q: q.clock.step * 10, p, freq=2, nsteps=4)
    assert d.N == 3 and d.i == 1 and d.data[0] == 0
    for s in range(1, 11):
        p.clock.step, p.clock.t = s, 0.1 * s
        M.increment([d])
    assert d.i == 6 and d["steps"] == [0, 2, 4, 6, 8, 10] and d["data"][-1] == 100
    assert len(d.data) >= 6 and d() == 100 and d[1] == 20


def test_round_sig_matches_julia_round_sigdigits():
    assert P._round_sig(42.8134) == 42.8 and P._round_sig(0.0123456) == 0.0123 and P._round_sig(123456.0) == 123000.0
    assert P._round_sig(0.0) == 0.0 and math.isnan(P._round_sig(float("nan")))


class _FakeFieldProb:
    """Duck-typed problem for the host-only checkpoint logic (mhdflows_jl_b200/io.py)."""

    def __init__(self, b=True, e=False):
        from mhdflows_jl_b200.problem import _Flag
        self.flag = _Flag(b, e)
        self.clock = _FakeClock()
        self.rank, self.nranks, self._real_shape = 0, 1, (4, 4, 4)
        rng = np.random.default_rng(0)
        names = ["bx", "by", "bz"] if e else ["ux", "uy", "uz"] + (["bx", "by", "bz"] if b else [])
        self.fields = {n: rng.standard_normal((4, 4, 4)).astype(np.float32) for n in names}
        self.which = []

    def get_real(self, f, which=0):
        self.which.append(which)
        return self.fields[f].copy()

    def set_real(self, f, arr):
        self.fields[f] = np.array(arr, dtype=np.float32)


@pytest.mark.parametrize("b,e", [(False, False), (True, False), (True, True)])
def test_savefile_restart_roundtrip(tmp_path, b, e):
    """savefile / Restart! (integrator.jl:208-288): dataset names of the reference, stale vars, clock.t restored only."""
    p = _FakeFieldProb(b, e)
    p.clock.t, p.clock.step = 1.25, 7
    path = M.savefile(p, 3, file_path_and_name=str(tmp_path / "run"))
    assert path.endswith("run_t_0003.h5") and set(p.which) == {M.STALE}
    d = M.readMHDFlows(path)
    want = {"time"} | ({"i_mag_field", "j_mag_field", "k_mag_field"} if b else set()) | (set() if e else {"i_velocity", "j_velocity", "k_velocity"})
    assert set(d) == want and float(d["time"]) == 1.25
    q = _FakeFieldProb(b, e)
    for k in q.fields:
        q.fields[k][...] = 0
    M.Restart(q, path)
    assert q.clock.t == 1.25 and q.clock.step == 0
    for k in p.fields:
        assert np.array_equal(q.fields[k], p.fields[k])


def test_a99_normalisation_matches_the_oracle_setup():
    """SetUpFk's amplitude A (pgen/A99ForceDriving.jl:104-105): the host mirror sums the shell weights plane by plane in
    Float64, the oracle evaluates the reference's array expressions in T."""
    import math

    from mhdflows_jl_b200.problem import _a99_integral, _Grid
    from oracle import forcing_oracle as FO
    from oracle import mhdflows_oracle as O
    for T, tol in ((np.float32, 2e-6), (np.float64, 1e-13)):
        op = O.Problem(nx=16, ny=32, nz=8, Lx=2 * math.pi, Ly=4 * math.pi, T=T, B_field=True)
        op.vars.usr_vars = FO.A99Vars(op.grid)
        A_ref = FO.SetUpFk(op, kf=3, P=2, sigma2=1.5)
        g = _Grid(16, 32, 8, 2 * math.pi, 4 * math.pi, 2 * math.pi, T)
        A = math.sqrt(2 * 3 * (g.Lx / g.dx) * (g.Ly / g.dy) * (g.Lz / g.dz) / _a99_integral(g, 3.0, 1.5) * (1 / g.dx / g.dy / g.dz))
        assert abs(A - A_ref) < tol * A_ref


def test_vp_fields_are_routed_to_the_library():
    """Problem(...; VP_method=true): params.χ and the U₀ / B₀ keywords of SetUpProblemIC! (IC.jl:93-106) end up in
    set_vp_field with the header's numbering (0 chi, 1..3 U0, 4..6 B0); B₀ is ignored without a magnetic field."""
    from mhdflows_jl_b200.problem import _Flag, _Params

    class Fake:
        def __init__(self, b):
            self.flag = _Flag(b, False, vp=True)
            self.sent = []
            self.params = _Params(self)

        def set_vp_field(self, name, arr):
            self.sent.append((P._VP_FIELDS[name], float(np.asarray(arr).ravel()[0])))

        def set_real(self, f, arr):
            self.sent.append((f, None))

    for b in (False, True):
        p = Fake(b)
        p.params.χ = np.full((2, 2, 2), 7.0)
        p.params.ν = 0.1                                   # ordinary members stay ordinary
        M.SetUpProblemIC(p, ux=np.ones((2, 2, 2)), U0y=np.full((2, 2, 2), 2.0), B0z=np.full((2, 2, 2), 3.0), **{"U₀x": np.full((2, 2, 2), 1.0)})
        want = [(0, 7.0), (2, 2.0)] + ([(6, 3.0)] if b else []) + [(1, 1.0), ("ux", None)]
        assert sorted(map(str, p.sent)) == sorted(map(str, want)) and p.params.ν == 0.1
    with pytest.raises(NotImplementedError):
        M.SetUpProblemIC(Fake(True), rho=np.ones((2, 2, 2)))


def test_cylindrical_mask_function():
    """Cylindrical_Mask_Function (utils/IC.jl:5-27): 0 inside the annulus R1 <= R <= R2, 1 in the solid, constant along z."""
    g = P._Grid(32, 32, 8, 2 * math.pi, 2 * math.pi, 2 * math.pi, np.float32)
    S = M.Cylindrical_Mask_Function(g)
    assert S.shape == (8, 32, 32) and S.dtype == np.float32 and set(np.unique(S)) == {0.0, 1.0}
    assert np.array_equal(S[0], S[5])
    for j in (0, 7, 16, 31):
        for i in (0, 3, 16, 30):
            R = math.sqrt(float(g.x[i]) ** 2 + float(g.y[j]) ** 2)
            assert S[0, j, i] == (0.0 if R <= 0.82 * math.pi else 1.0)
    S2 = M.Cylindrical_Mask_Function(g, R1=1.0, **{"R₂": 2.0})
    assert S2[0, 16, 16] == 1.0 and S2[0, 16, 16 + 8] == 0.0     # R = 0 is solid, R = 8 dx = 1.57 is fluid
