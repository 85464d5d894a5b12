"""CPU oracle for the MHDFlows.jl hot path -- TEST INFRASTRUCTURE ONLY.

Literal NumPy/SciPy(pocketfft) restatement of the reference's 3D periodic
pseudospectral right-hand side + RK4/LSRK54 time step (HD / MHD / EMHD),
its CFL rule, dashboard diagnostics, helicities, shell spectrum and the
random-phase solenoidal initial condition.  "Literal" means: same number and
order of FFTs (36 / 24 / 51 per RHS evaluation), same in-place dealias of the
stage input, same `rfft(irfft(sol))` diffusion operand, same stale `vars`.

PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures
(SURVEY.md section 4 / 8c), and Julia is not installed here, so this oracle
cannot be checked against reference outputs.  Its fidelity rests on (i)
line-by-line correspondence with the files cited in each docstring, (ii) the
physics known-answer tests in tests/test_oracle_known_answers.py, (iii) the
README timing ballpark.  Semantics of the un-vendored FourierFlows.jl v0.10.1
(Manifest.toml:116-120) are restated from its published algorithm.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module.  The product path
(mhdflows_jl_b200) never does.

Array convention: Julia column-major (nx, ny, nz) == NumPy C-order (nz, ny, nx).
Spectral fields: Julia (nkr, nl, nm, Nfield) == NumPy (Nfield, nm, nl, nkr).
All file:line citations are relative to /root/reference/.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field
from fractions import Fraction

import numpy as np
import scipy.fft as sfft

_WORKERS = int(os.environ.get("MHDF_ORACLE_WORKERS", os.cpu_count() or 1))


def set_workers(n: int) -> None:
    global _WORKERS
    _WORKERS = int(n)


# --------------------------------------------------------------------------
# Grid  (FourierFlows.ThreeDGrid; in-repo mirror src/utils/utils.jl:42-98)
# --------------------------------------------------------------------------
def aliased_range(nk: int, aliased_fraction: float = 1 / 3):
    """FourierFlows.getaliasedwavenumbers: 1-based inclusive (iL, iR).
    Evaluated in Float64 exactly like the Julia expression."""
    L = (1 - aliased_fraction) / 2
    R = (1 + aliased_fraction) / 2
    iL = math.floor(L * nk) + 1
    iR = math.ceil(R * nk)
    return iL, iR


class Grid:
    """ThreeDGrid restatement (src/utils/utils.jl:42-98 is the in-repo copy).

    kr, l, m are built in Float64 then converted to T; Krsq / invKrsq are
    computed in T from the T vectors; invKrsq[0,0,0] = 0."""

    def __init__(self, nx, ny=None, nz=None, Lx=2 * math.pi, Ly=None, Lz=None, T=np.float32):
        ny = nx if ny is None else ny
        nz = nx if nz is None else nz
        Ly = Lx if Ly is None else Ly
        Lz = Lx if Lz is None else Lz
        self.nx, self.ny, self.nz = nx, ny, nz
        self.Lx, self.Ly, self.Lz = Lx, Ly, Lz
        self.T = np.dtype(T).type
        self.CT = np.complex64 if self.T is np.float32 else np.complex128
        self.nkr = nx // 2 + 1
        self.nl, self.nm = ny, nz
        self.dx, self.dy, self.dz = Lx / nx, Ly / ny, Lz / nz
        T_ = self.T
        # x = range(T(x0), step=T(dx), length=nx), x0 = -L/2
        self.x = (T_(-Lx / 2) + T_(self.dx) * np.arange(nx)).astype(T_)
        self.y = (T_(-Ly / 2) + T_(self.dy) * np.arange(ny)).astype(T_)
        self.z = (T_(-Lz / 2) + T_(self.dz) * np.arange(nz)).astype(T_)
        # rfftfreq(nx, 2pi/Lx*nx), fftfreq(n, 2pi/L*n): Nyquist negative
        self.kr = (np.arange(self.nkr) * (2 * math.pi / Lx)).astype(T_).reshape(1, 1, -1)
        self.l = (np.fft.fftfreq(ny, 1.0 / ny) * (2 * math.pi / Ly)).astype(T_).reshape(1, -1, 1)
        self.m = (np.fft.fftfreq(nz, 1.0 / nz) * (2 * math.pi / Lz)).astype(T_).reshape(-1, 1, 1)
        self.Krsq = (self.kr ** 2 + self.l ** 2 + self.m ** 2).astype(T_)
        with np.errstate(divide="ignore"):
            self.invKrsq = (T_(1) / self.Krsq).astype(T_)
        self.invKrsq[0, 0, 0] = 0
        iL, _ = aliased_range(nx)
        self.kralias = slice(iL - 1, self.nkr)  # iL:nkr (1-based)
        iL, iR = aliased_range(ny)
        self.lalias = slice(iL - 1, iR)
        iL, iR = aliased_range(nz)
        self.malias = slice(iL - 1, iR)

    # unnormalised forward r2c over dims 1:3 (mul!(yh, rfftplan, y))
    def rfft(self, f):
        return sfft.rfftn(f, axes=(0, 1, 2), workers=_WORKERS).astype(self.CT, copy=False)

    # ldiv!(y, rfftplan, yh): c2r scaled by 1/(nx ny nz); x axis is the c2r axis
    def irfft(self, fh):
        return sfft.irfftn(fh, s=(self.nz, self.ny, self.nx), axes=(0, 1, 2),
                           workers=_WORKERS).astype(self.T, copy=False)

    def dealias(self, fh):
        """FourierFlows.dealias!(fh, grid::ThreeDGrid): three strided zero fills, in place
        (call sites pgen.jl:155,166,174)."""
        fh[..., self.kralias] = 0
        fh[..., self.lalias, :] = 0
        fh[..., self.malias, :, :] = 0
        return fh

    def retained_mask(self):
        msk = np.ones((self.nm, self.nl, self.nkr), dtype=bool)
        msk[..., self.kralias] = False
        msk[:, self.lalias, :] = False
        msk[self.malias, :, :] = False
        return msk


# --------------------------------------------------------------------------
# vars / params / clock   (src/Structure/datastructure.jl:5-35,58-108)
# --------------------------------------------------------------------------
@dataclass
class Clock:  # FourierFlows.Clock{T}(dt, t, step)   Problems.jl:120
    dt: float
    t: float = 0.0
    step: int = 0


@dataclass
class Flag:  # Problems.jl:68-79
    b: bool = False
    e: bool = False
    vp: bool = False
    c: bool = False
    s: bool = False


class Vars:
    """MVars / HVars / EMVars: real-space fields + two scratch arrays
    (datastructure.jl:5-35, MHDVars.jl:1-49, HDVars.jl:1-18)."""

    def __init__(self, grid: Grid, B: bool, E: bool):
        z = lambda: np.zeros((grid.nz, grid.ny, grid.nx), dtype=grid.T)
        if E:
            self.bx, self.by, self.bz = z(), z(), z()
            self.curlBx, self.curlBy, self.curlBz = z(), z(), z()  # nabla x B (i, j, k)
        else:
            self.ux, self.uy, self.uz = z(), z(), z()
            if B:
                self.bx, self.by, self.bz = z(), z(), z()
        self.nonlin1 = z()
        self.nonlinh1 = np.zeros((grid.nm, grid.nl, grid.nkr), dtype=grid.CT)


@dataclass
class Params:
    """MHDParams / HDParams / EMHDParams (MHDParams.jl:42-94, HDParams.jl:30-51).
    nu/eta are stored as given (Float64) -> mixed-precision broadcast.
    0-based field indices."""
    nu: float = 0.0
    eta: float = 0.0
    n_nu: int = 0
    n_eta: int = 0  # never forwarded by Problem (pgen.jl:114-116) -> always 0
    calcF: object = None
    vp: object = None   # VPParams when Problem(VP_method=True): chi, U0x.., B0x.. (datastructure.jl:80-81,94-95)
    ux_ind: int = 0
    uy_ind: int = 1
    uz_ind: int = 2
    bx_ind: int = 3
    by_ind: int = 4
    bz_ind: int = 5


# --------------------------------------------------------------------------
# RHS -- HD   (src/Solver/HDSolver.jl:25-108)
# --------------------------------------------------------------------------
def _delta(a, b):
    return 1 if a == b else 0


class VPParams:
    """The volume-penalisation members of MHDParams_VP / HDParams_VP (datastructure.jl:80-81,94-95): real fields chi,
    U0x, U0y, U0z (, B0x, B0y, B0z), all zero after construction; `clock` is the problem's clock (VP_*Update! receive it)."""

    def __init__(self, grid, B, clock):
        z = lambda: np.zeros((grid.nz, grid.ny, grid.nx), dtype=grid.T)
        self.chi, self.U0x, self.U0y, self.U0z = z(), z(), z(), z()
        if B:
            self.B0x, self.B0y, self.B0z = z(), z(), z()
        self.clock = clock


def _vp_update(dfdt, ka_kinv2, a, fs, Ws, vars, params, grid):
    """VPSolver.VP_UiUpdate! / VP_BiUpdate! (VPSolver.jl:21-59): for j = x,y,z:
    tmp = chi/eta*(f_j - W_j), eta = clock.dt*13/7;  df_a/dt += -(delta(a,j) - k_j*(k_a k^-2)) * rfft(tmp)."""
    T = grid.T
    ks = (grid.kr, grid.l, grid.m)
    chi = params.vp.chi
    eta = T(params.vp.clock.dt) * T(13) / T(7)
    for j in range(3):
        vars.nonlin1[...] = chi / eta * (fs[j] - Ws[j])
        vars.nonlinh1[...] = grid.rfft(vars.nonlin1)
        dfdt += -(T(_delta(a, j)) - ks[j] * ka_kinv2) * vars.nonlinh1


def _hd_Ui_update(N, sol, vars, params, grid, a):
    """HDSolver.UiUpdate! (HDSolver.jl:25-93); a in {0,1,2}."""
    T, CT = grid.T, grid.CT
    ks = (grid.kr, grid.l, grid.m)
    us = (vars.ux, vars.uy, vars.uz)
    ka = ks[a]
    kinv2 = grid.invKrsq
    dudt = N[(params.ux_ind, params.uy_ind, params.uz_ind)[a]]
    dudt *= 0                                                   # :55
    for i in range(3):
        for j in range(3):
            if i <= j:
                vars.nonlin1[...] = us[i] * us[j]               # :62
                vars.nonlinh1[...] = grid.rfft(vars.nonlin1)    # :65
                uuh = vars.nonlinh1
                dudt += (CT(-1j) * ks[i] * (T(_delta(a, j)) - ka * ks[j] * kinv2)) * uuh       # :68
                if i != j:
                    dudt += (CT(-1j) * ks[j] * (T(_delta(a, i)) - ka * ks[i] * kinv2)) * uuh   # :70
    if params.vp is not None:                                   # :76-79
        vp = params.vp
        _vp_update(dudt, ka * kinv2, a, us, (vp.U0x, vp.U0y, vp.U0z), vars, params, grid)
    uh = vars.nonlinh1
    uh[...] = grid.rfft(us[a])                                  # :84
    # -Krsq*nu*uh with nu Float64 -> evaluated in Float64, rounded on store
    dudt += (-(grid.Krsq.astype(np.float64)) * params.nu * uh).astype(CT)   # :85
    if params.n_nu > 1:                                         # :88-90
        dudt += (-(grid.Krsq.astype(np.float64) ** params.n_nu) * params.nu * uh).astype(CT)


def hd_calcN_advection(N, sol, vars, params, grid):
    """HDSolver.HDcalcN_advection! (HDSolver.jl:95-108): 3 c2r + 3*(6+1) r2c = 24 FFTs."""
    vars.ux[...] = grid.irfft(sol[params.ux_ind].copy())
    vars.uy[...] = grid.irfft(sol[params.uy_ind].copy())
    vars.uz[...] = grid.irfft(sol[params.uz_ind].copy())
    for a in range(3):
        _hd_Ui_update(N, sol, vars, params, grid, a)


def HDcalcN(N, sol, t, clock, vars, params, grid):
    """pgen.jl:173-181 -- note forcing is applied BEFORE advection, which zeroes N
    (HDSolver.jl:55), so HD forcing is lost."""
    grid.dealias(sol)
    if params.calcF is not None:
        params.calcF(N, sol, t, clock, vars, params, grid)
    hd_calcN_advection(N, sol, vars, params, grid)


# --------------------------------------------------------------------------
# RHS -- MHD   (src/Solver/MHDSolver.jl:27-177,330-351)
# --------------------------------------------------------------------------
def _mhd_Ui_update(N, sol, vars, params, grid, a):
    """MHDSolver.UiUpdate! (MHDSolver.jl:27-103)."""
    T, CT = grid.T, grid.CT
    ks = (grid.kr, grid.l, grid.m)
    us = (vars.ux, vars.uy, vars.uz)
    bs = (vars.bx, vars.by, vars.bz)
    ka = ks[a]
    kinv2 = grid.invKrsq
    dudt = N[(params.ux_ind, params.uy_ind, params.uz_ind)[a]]
    dudt *= 0                                                   # :62
    for i in range(3):
        for j in range(3):
            if j >= i:
                vars.nonlin1 *= 0                               # :67
                vars.nonlinh1 *= 0                              # :68
                vars.nonlin1[...] = bs[i] * bs[j] - us[i] * us[j]        # :73
                vars.nonlinh1[...] = grid.rfft(vars.nonlin1)             # :74
                Th = vars.nonlinh1
                dudt += (CT(1j) * ks[i] * (T(_delta(a, j)) - ka * ks[j] * kinv2)) * Th       # :77
                if i != j:
                    dudt += (CT(1j) * ks[j] * (T(_delta(a, i)) - ka * ks[i] * kinv2)) * Th   # :79
    if params.vp is not None:                                   # :85-88
        vp = params.vp
        _vp_update(dudt, ka * kinv2, a, us, (vp.U0x, vp.U0y, vp.U0z), vars, params, grid)
    uh = vars.nonlinh1
    uh[...] = grid.rfft(us[a])                                  # :93
    dudt += (-(grid.Krsq.astype(np.float64)) * params.nu * uh).astype(CT)   # :94
    if params.n_nu > 1:                                         # :97-99 (adds ON TOP of :94)
        dudt += (-(grid.Krsq.astype(np.float64) ** params.n_nu) * params.nu * uh).astype(CT)


def _mhd_Bi_update(N, sol, vars, params, grid, a):
    """MHDSolver.BiUpdate! (MHDSolver.jl:106-177)."""
    CT = grid.CT
    ks = (grid.kr, grid.l, grid.m)
    us = (vars.ux, vars.uy, vars.uz)
    bs = (vars.bx, vars.by, vars.bz)
    dbdt = N[(params.bx_ind, params.by_ind, params.bz_ind)[a]]
    dbdt *= 0                                                   # :143
    for j in range(3):
        if a != j:
            vars.nonlin1[...] = us[a] * bs[j] - bs[a] * us[j]   # :150
            vars.nonlinh1[...] = grid.rfft(vars.nonlin1)        # :152
            dbdt += (CT(1j) * ks[j]) * vars.nonlinh1            # :155
    if params.vp is not None:                                   # :160-163
        vp = params.vp
        _vp_update(dbdt, ks[a] * grid.invKrsq, a, bs, (vp.B0x, vp.B0y, vp.B0z), vars, params, grid)
    bh = vars.nonlinh1
    bh[...] = grid.rfft(bs[a])                                  # :167
    dbdt += (-(grid.Krsq.astype(np.float64)) * params.eta * bh).astype(CT)  # :168
    if params.n_eta > 1:                                        # :171-173 (never active)
        dbdt += (-(grid.Krsq.astype(np.float64) ** params.n_eta) * params.eta * bh).astype(CT)


def mhd_calcN_advection(N, sol, vars, params, grid):
    """MHDSolver.MHDcalcN_advection! (MHDSolver.jl:330-351): 6 + 3*7 + 3*3 = 36 FFTs."""
    for name, ind in (("ux", params.ux_ind), ("uy", params.uy_ind), ("uz", params.uz_ind),
                      ("bx", params.bx_ind), ("by", params.by_ind), ("bz", params.bz_ind)):
        getattr(vars, name)[...] = grid.irfft(sol[ind].copy())   # :333-338
    for a in range(3):
        _mhd_Ui_update(N, sol, vars, params, grid, a)            # :341-343
    for a in range(3):
        _mhd_Bi_update(N, sol, vars, params, grid, a)            # :346-348


def MHDcalcN(N, sol, t, clock, vars, params, grid):
    """pgen.jl:153-162."""
    grid.dealias(sol)
    mhd_calcN_advection(N, sol, vars, params, grid)
    if params.calcF is not None:
        params.calcF(N, sol, t, clock, vars, params, grid)


# --------------------------------------------------------------------------
# RHS -- EMHD   (src/Solver/MHDSolver.jl:183-328)
# --------------------------------------------------------------------------
def _get_curlB(sol, vars, params, grid):
    """MHDSolver.Get(nabla x B)! "way 2" (MHDSolver.jl:273-312): 3 c2r."""
    CT = grid.CT
    k1, k2, k3 = grid.kr, grid.l, grid.m
    B1h, B2h, B3h = sol[params.bx_ind], sol[params.by_ind], sol[params.bz_ind]
    CBh = vars.nonlinh1
    CBh[...] = CT(1j) * (k2 * B3h - k3 * B2h)
    vars.curlBx[...] = grid.irfft(CBh)
    CBh[...] = CT(1j) * (k3 * B1h - k1 * B3h)
    vars.curlBy[...] = grid.irfft(CBh)
    CBh[...] = CT(1j) * (k1 * B2h - k2 * B1h)
    vars.curlBz[...] = grid.irfft(CBh)


def _emhd_Bi_update(N, sol, vars, params, grid, a):
    """MHDSolver.EMHD_BiUpdate! (MHDSolver.jl:183-270): 15 FFTs.
    dB_i/dt = F[A_j d_j B_i] - F[b_j d_j A_i], A = curl B, d_i = 1;
    b_j is the STALE vars.b* (refreshed only at the end of EMHDcalcN_advection!)."""
    CT = grid.CT
    ks = (grid.kr, grid.l, grid.m)
    A = (vars.curlBx, vars.curlBy, vars.curlBz)
    bs = (vars.bx, vars.by, vars.bz)
    ind = (params.bx_ind, params.by_ind, params.bz_ind)[a]
    Ai = A[a]
    bih = sol[ind]
    dBdt = N[ind]
    dBdt *= 0                                                   # :240
    for j in range(3):
        # (B . grad) A_i
        vars.nonlinh1[...] = 0
        vars.nonlinh1[...] = grid.rfft(Ai)                      # :245
        vars.nonlinh1[...] = CT(1j) * ks[j] * vars.nonlinh1     # :246
        vars.nonlin1[...] = grid.irfft(vars.nonlinh1.copy())    # :247
        vars.nonlin1[...] = bs[j] * vars.nonlin1                # :249
        vars.nonlinh1[...] = grid.rfft(vars.nonlin1)            # :251
        dBdt -= vars.nonlinh1                                   # :253
        # (A . grad) B_i
        vars.nonlinh1[...] = CT(1j) * ks[j] * bih               # :257
        vars.nonlin1[...] = grid.irfft(vars.nonlinh1.copy())    # :258
        vars.nonlin1[...] = A[j] * vars.nonlin1                 # :260
        vars.nonlinh1[...] = grid.rfft(vars.nonlin1)            # :262
        dBdt += vars.nonlinh1                                   # :264


def emhd_calcN_advection(N, sol, vars, params, grid):
    """MHDSolver.EMHDcalcN_advection! (MHDSolver.jl:314-328): 3 + 45 + 3 = 51 FFTs."""
    _get_curlB(sol, vars, params, grid)
    for a in range(3):
        _emhd_Bi_update(N, sol, vars, params, grid, a)
    vars.bx[...] = grid.irfft(sol[params.bx_ind].copy())        # :323-325
    vars.by[...] = grid.irfft(sol[params.by_ind].copy())
    vars.bz[...] = grid.irfft(sol[params.bz_ind].copy())


def EMHDcalcN(N, sol, t, clock, vars, params, grid):
    """pgen.jl:164-171: no forcing, no resistive term on this path."""
    grid.dealias(sol)
    emhd_calcN_advection(N, sol, vars, params, grid)


# --------------------------------------------------------------------------
# Time steppers (FourierFlows RK4TimeStepper / LSRK54TimeStepper; RK4 mirror
# src/DyeModule.jl:62-91; LSRK54 = Carpenter & Kennedy 1994)
# --------------------------------------------------------------------------
LSRK54_A = (Fraction(0), Fraction(-567301805773, 1357537059087), Fraction(-2404267990393, 2016746695238),
            Fraction(-3550918686646, 2091501179385), Fraction(-1275806237668, 842570457699))
LSRK54_B = (Fraction(1432997174477, 9575080441755), Fraction(5161836677717, 13612068292357),
            Fraction(1720146321549, 2090206949498), Fraction(3134564353537, 4481467310338),
            Fraction(2277821191437, 14882151754819))
LSRK54_C = (Fraction(0), Fraction(1432997174477, 9575080441755), Fraction(2526269341429, 6820363962896),
            Fraction(2006345519317, 3224310063776), Fraction(2802321613138, 2924317926251))


class RK4TimeStepper:
    def __init__(self, sol_like):
        self.sol1 = np.zeros_like(sol_like)
        self.RHS = [np.zeros_like(sol_like) for _ in range(4)]


class LSRK54TimeStepper:
    def __init__(self, sol_like, T):
        self.S2 = np.zeros_like(sol_like)
        self.RHS = np.zeros_like(sol_like)
        self.A = [T(float(a)) for a in LSRK54_A]
        self.B = [T(float(b)) for b in LSRK54_B]
        self.C = [T(float(c)) for c in LSRK54_C]


class HM89TimeStepper:
    """timestepper/HM89.jl:7-23: five work arrays F0, F1, B0, B1, Bn and c = (1//3, 15//16, 8//15)."""

    def __init__(self, sol_like, T):
        self.F0, self.F1, self.B0, self.B1, self.Bn = (np.zeros_like(sol_like) for _ in range(5))
        self.c = tuple(T(float(c)) for c in (Fraction(1, 3), Fraction(15, 16), Fraction(8, 15)))   # Rational -> T in the broadcasts
        self.iters = 0          # fixed-point iterations of the last step (bookkeeping of this restatement)
        self.eps = 0.0          # last error norm
        self.dealias_vars = False   # True: the closing ldiv! reads a dealiased copy of sol (what a pruned-spectrum
                                    # implementation holds); False = the reference, literally


def DivFreeCorrection(sol, vars, params, grid):
    """timestepper/HM89.jl:105-139: Phi = -i (k . B^) / k^2, B^_i -= i k_i Phi."""
    CT = grid.CT
    ki, kj, kk = grid.kr, grid.l, grid.m
    vars.nonlin1 *= 0
    vars.nonlinh1 *= 0
    S = vars.nonlinh1
    bxh, byh, bzh = sol[params.bx_ind], sol[params.by_ind], sol[params.bz_ind]
    S[...] = CT(-1j) * (ki * bxh + kj * byh + kk * bzh)          # :129
    S[...] = S * grid.invKrsq                                    # :130
    bxh -= CT(1j) * ki * S                                       # :133-135
    byh -= CT(1j) * kj * S
    bzh -= CT(1j) * kk * S


def LSRK3substeps(sol, clock, ts, calcN, vars, params, grid):
    """timestepper/HM89.jl:141-167, literally: the first stage multiplies by dt twice (:157-158), 5/9 and 153/128 are
    Float64 constants (the broadcast runs in ComplexF64 and rounds on the store)."""
    T = grid.T
    t, dt, c = clock.t, T(clock.dt), ts.c
    calcN(ts.F0, sol, t + dt, clock, vars, params, grid)
    ts.F0 *= dt
    sol += ts.F0 * c[0] * dt
    calcN(ts.F1, sol, t + dt, clock, vars, params, grid)
    ts.F1 *= dt
    ts.F1 -= np.float64(5 / 9) * ts.F0
    sol += c[1] * ts.F1
    calcN(ts.F0, sol, t + dt, clock, vars, params, grid)
    ts.F0 *= dt
    ts.F0 -= np.float64(153 / 128) * ts.F1
    sol += c[2] * ts.F0


def RK3linearterm(sol, ts, clock, vars, params, grid):
    """timestepper/HM89.jl:169-199: the same three stages on F = calcF!(...) - eta k^2 sol.  params.calcF! = nothingfunction
    leaves F0 / F1 as the fixed-point loop left them (F0 = the last B^n - B^1, F1 = the last curl(J x B)): they ARE read."""
    T = grid.T
    t, dt, c = clock.t, T(clock.dt), ts.c
    k2, eta = grid.Krsq, np.float64(params.eta)
    calcF = params.calcF if params.calcF is not None else (lambda *a: None)
    calcF(ts.F0, sol, t + dt, clock, vars, params, grid)
    ts.F0 -= eta * k2 * sol
    ts.F0 *= dt
    sol += ts.F0 * c[0] * dt
    calcF(ts.F1, sol, t + dt, clock, vars, params, grid)
    ts.F1 -= eta * k2 * sol
    ts.F1 *= dt
    ts.F1 -= np.float64(5 / 9) * ts.F0
    sol += c[1] * ts.F1
    calcF(ts.F0, sol, t + dt, clock, vars, params, grid)
    ts.F0 -= eta * k2 * sol
    ts.F0 *= dt
    ts.F0 -= np.float64(153 / 128) * ts.F1
    sol += c[2] * ts.F0


def HM89substeps(sol, clock, ts, calcN, vars, params, grid):
    """timestepper/HM89.jl:35-103: LSRK3 predictor, divergence cleaning, fixed-point iteration of the implicit midpoint rule
    on the Hall term (error = max |B^n - B^1| in real space, threshold 5e-4, written into vars.b*: the NEXT evaluation of
    the loop reads that difference as its stale b), then the resistive term with the explicit three-stage scheme."""
    T = grid.T
    t, dt = clock.t, T(clock.dt)
    B0, B1, Bn = ts.B0, ts.B1, ts.Bn
    dBh, NL = ts.F0, ts.F1
    B0[...] = sol                                                # :54
    LSRK3substeps(sol, clock, ts, calcN, vars, params, grid)
    DivFreeCorrection(sol, vars, params, grid)
    B1[...] = sol
    grid.dealias(B1)
    B_half = sol
    eps, err = 1.0, 5e-4
    ts.iters = 0
    while eps > err:
        B_half[...] = (B0 + B1) * 0.5                            # :67
        calcN(NL, B_half, t, clock, vars, params, grid)
        Bn[...] = B0 + dt * NL                                   # :71
        grid.dealias(Bn)
        dBh[...] = Bn - B1                                       # :75
        vars.bx[...] = grid.irfft(dBh[0].copy())                 # :76-78 (dBx, dBy, dBz ARE vars.bx, by, bz)
        vars.by[...] = grid.irfft(dBh[1].copy())
        vars.bz[...] = grid.irfft(dBh[2].copy())
        eps = float(np.max(np.sqrt(vars.bx * vars.bx + vars.by * vars.by + vars.bz * vars.bz)))   # :79
        B1[...] = Bn
        ts.iters += 1
    ts.eps = eps
    sol[...] = B1                                                # :86
    RK3linearterm(sol, ts, clock, vars, params, grid)
    DivFreeCorrection(sol, vars, params, grid)
    src = grid.dealias(sol.copy()) if ts.dealias_vars else sol
    vars.bx[...] = grid.irfft(src[params.bx_ind].copy())         # :91-93
    vars.by[...] = grid.irfft(src[params.by_ind].copy())
    vars.bz[...] = grid.irfft(src[params.bz_ind].copy())


def stepforward(prob):
    """timestepper/timestepper.jl:4-6 -> FourierFlows.stepforward!"""
    sol, clock, ts, vars, params, grid = prob.sol, prob.clock, prob.timestepper, prob.vars, prob.params, prob.grid
    calcN = prob.calcN
    T = grid.T
    dt = T(clock.dt)
    t = clock.t
    if isinstance(ts, HM89TimeStepper):                            # HM89.jl:25-33: clock.t += clock.dt (in T), step += 1
        HM89substeps(sol, clock, ts, calcN, vars, params, grid)
    elif isinstance(ts, RK4TimeStepper):
        R1, R2, R3, R4 = ts.RHS
        calcN(R1, sol, t, clock, vars, params, grid)               # L = 0: addlinearterm! is a no-op
        ts.sol1[...] = sol + (dt / T(2)) * R1
        calcN(R2, ts.sol1, t + dt / 2, clock, vars, params, grid)
        ts.sol1[...] = sol + (dt / T(2)) * R2
        calcN(R3, ts.sol1, t + dt / 2, clock, vars, params, grid)
        ts.sol1[...] = sol + dt * R3
        calcN(R4, ts.sol1, t + dt, clock, vars, params, grid)
        sol += dt * (R1 / T(6) + R2 / T(3) + R3 / T(3) + R4 / T(6))   # DyeModule.jl:62-66
    else:
        ts.S2[...] = 0
        for i in range(5):
            calcN(ts.RHS, sol, t + ts.C[i] * dt, clock, vars, params, grid)
            ts.S2[...] = ts.A[i] * ts.S2 + dt * ts.RHS
            sol += ts.B[i] * ts.S2
    clock.t = float(T(T(clock.t) + dt))   # clock.t :: T
    clock.step += 1


# --------------------------------------------------------------------------
# Problem  (src/pgen.jl:64-150, src/Problems.jl:12-18,118-140)
# --------------------------------------------------------------------------
class Problem:
    def __init__(self, nx=64, ny=None, nz=None, Lx=2 * math.pi, Ly=None, Lz=None, dt=0.0,
                 nu=0.0, n_nu=0, eta=0.0, n_eta=0, B_field=False, EMHD=False,
                 stepper="RK4", calcF=None, T=np.float32, aliased_fraction=1 / 3, VP_method=False):
        # aliased_fraction is accepted but NOT forwarded to the grid (pgen.jl:107)
        self.grid = Grid(nx, ny, nz, Lx, Ly, Lz, T)
        self.flag = Flag(b=B_field, e=EMHD)
        # SetVars: E wins over B (datastructure.jl:58-67); SetParams: EMHD params only when
        # B && E (datastructure.jl:78-88)
        self.vars = Vars(self.grid, B_field, EMHD)
        if EMHD and not B_field:
            raise ValueError("EMHD requires B_field=true (datastructure.jl:78-88)")
        if EMHD:
            self.params = Params(eta=eta, n_eta=0, calcF=calcF, bx_ind=0, by_ind=1, bz_ind=2)
            self.Nl, self.calcN = 3, EMHDcalcN
        elif B_field:
            self.params = Params(nu=nu, eta=eta, n_nu=n_nu, n_eta=0, calcF=calcF)  # n_eta dropped: pgen.jl:114-116
            self.Nl, self.calcN = 6, MHDcalcN
        else:
            self.params = Params(nu=nu, n_nu=n_nu, calcF=calcF)
            self.Nl, self.calcN = 3, HDcalcN
        g = self.grid
        self.sol = np.zeros((self.Nl, g.nm, g.nl, g.nkr), dtype=g.CT)
        self.clock = Clock(dt=float(g.T(dt)))
        self.flag.vp = bool(VP_method)
        if VP_method:
            if EMHD:
                raise ValueError("VP_method: the EMHD equation has no volume-penalisation terms")
            self.params.vp = VPParams(g, B_field, self.clock)
        if stepper == "RK4":
            self.timestepper = RK4TimeStepper(self.sol)
        elif stepper == "LSRK54":
            self.timestepper = LSRK54TimeStepper(self.sol, g.T)
        elif stepper == "HM89" and EMHD:                           # Problems.jl:124-126: only with EFlag
            self.timestepper = HM89TimeStepper(self.sol, g.T)
        else:
            raise ValueError(f"stepper {stepper!r} not in scope (RK4, LSRK54; HM89 with EMHD)")
        self.stepper = stepper


def SetUpProblemIC(prob, ux=None, uy=None, uz=None, bx=None, by=None, bz=None):
    """IC.jl:41-109: copy into vars, rfft into sol; no dealias, no projection;
    velocity skipped for EMHD (:69)."""
    g, v, p = prob.grid, prob.vars, prob.params
    if not prob.flag.e:
        for arr, name, ind in ((ux, "ux", p.ux_ind), (uy, "uy", p.uy_ind), (uz, "uz", p.uz_ind)):
            if arr is not None:
                getattr(v, name)[...] = arr
                prob.sol[ind] = g.rfft(getattr(v, name))
    if prob.flag.b:
        for arr, name, ind in ((bx, "bx", p.bx_ind), (by, "by", p.by_ind), (bz, "bz", p.bz_ind)):
            if arr is not None:
                getattr(v, name)[...] = arr
                prob.sol[ind] = g.rfft(getattr(v, name))


# --------------------------------------------------------------------------
# Driver loop, CFL, diagnostics  (src/integrator.jl:31-198, utils/UserInterface.jl:29,65-86)
# --------------------------------------------------------------------------
def getCFL(prob, t_diff, Coef=0.3):
    """integrator.jl:158-198.  Reads the (stale) real-space vars."""
    v, g = prob.vars, prob.grid
    sqmax = lambda A: float(np.max(A * A))
    if prob.flag.e:
        vmax = math.sqrt(max(sqmax(v.curlBx), sqmax(v.curlBy), sqmax(v.curlBz)))
    else:
        vmax = math.sqrt(max(sqmax(v.ux), sqmax(v.uy), sqmax(v.uz)))
    if prob.flag.b:
        vamax = math.sqrt(max(sqmax(v.bx), sqmax(v.by), sqmax(v.bz)))
        vmax = max(vmax, vamax)
    dl = min(g.dx, g.dy, g.dz)
    if prob.flag.e:
        dl = dl ** 2
    with np.errstate(divide="ignore"):
        dt = min(Coef * dl / vmax if vmax > 0 else math.inf, t_diff)
    prob.clock.dt = float(g.T(dt))
    return prob.clock.dt


def round_sig(x, sig=3):
    if x == 0 or not math.isfinite(x):
        return x
    return round(x, sig - int(math.floor(math.log10(abs(x)))) - 1)


def ProbDiagnostic(prob, rounded=True):
    """UserInterface.jl:65-86: KE = round(sum(u^2) * dV, sigdigits=3) (no 1/2), from stale vars."""
    g, v = prob.grid, prob.vars
    dV = float(g.x[1] - g.x[0]) * float(g.y[1] - g.y[0]) * float(g.z[1] - g.z[0])
    rs = round_sig if rounded else (lambda x: x)
    out = []
    if not prob.flag.e:
        KE = float(np.sum(v.ux.astype(np.float64) ** 2 + v.uy.astype(np.float64) ** 2 + v.uz.astype(np.float64) ** 2)) * dV
        if math.isnan(KE):
            raise FloatingPointError("detected NaN! Quit the simulation right now.")
        out.append(rs(KE))
    if prob.flag.b:
        ME = float(np.sum(v.bx.astype(np.float64) ** 2 + v.by.astype(np.float64) ** 2 + v.bz.astype(np.float64) ** 2)) * dV
        if math.isnan(ME):
            raise FloatingPointError("detected NaN! Quit the simulation right now.")
        out.append(rs(ME))
    return tuple(out) if len(out) > 1 else out[0]


def TimeIntegrator(prob, t0, N0, usr_dt=0.0, CFL_Coef=0.25, diags=(), on_step=None):
    """integrator.jl:31-156 (loop + CFL only).  Runs while N0 >= step && t0 >= t, i.e. N0+1 steps."""
    p, g = prob.params, prob.grid
    if prob.flag.b:
        vi = p.eta if prob.flag.e else max(p.nu, p.eta)
        nv = p.n_eta if prob.flag.e else max(p.n_nu, p.n_eta)
    else:
        vi, nv = p.nu, p.n_nu
    dl = min(g.Lx / g.nx, g.Ly / g.ny, g.Lz / g.nz)
    with np.errstate(divide="ignore"):
        if vi == 0:
            t_diff = math.inf
        else:
            t_diff = CFL_Coef * dl ** nv / vi if nv > 1 else CFL_Coef * dl ** 2 / vi
    prob.clock.step = 0
    if usr_dt != 0.0:
        prob.clock.dt = float(g.T(usr_dt))

    def vp_corrections():                                       # integrator.jl:85-88, 118-122
        if prob.flag.vp:
            from . import forcing_oracle as FO
            FO.DivVCorrection(prob)
            if prob.flag.b:
                FO.DivBCorrection(prob)

    vp_corrections()
    while N0 >= prob.clock.step and t0 >= prob.clock.t:
        if usr_dt == 0.0:
            getCFL(prob, t_diff, Coef=CFL_Coef)
        stepforward(prob)
        for d in diags:
            d.increment()
        vp_corrections()
        if on_step is not None:
            on_step(prob)


class Diagnostic:
    """DiagnosticWrapper.jl:14-105."""

    def __init__(self, calc, prob, freq=1, nsteps=100, ndata=None):
        ndata = math.ceil((nsteps + 1) / freq) if ndata is None else ndata
        self.calc, self.prob, self.freq, self.N = calc, prob, freq, ndata
        self.data = [None] * ndata
        self.t = [0.0] * ndata
        self.steps = [0] * ndata
        self.data[0], self.t[0], self.steps[0] = calc(prob), prob.clock.t, prob.clock.step
        self.i = 1

    def update(self, i):
        if i > len(self.steps):
            self.data += [None] * self.N
            self.t += [0.0] * self.N
            self.steps += [0] * self.N
        self.data[i - 1], self.t[i - 1], self.steps[i - 1] = self.calc(self.prob), self.prob.clock.t, self.prob.clock.step
        self.i = i

    def increment(self):
        if self.prob.clock.step % self.freq == 0:
            self.update(self.i + 1)


# --------------------------------------------------------------------------
# Analysis: curl, helicities, spectrum (utils/VectorCalculus.jl:24-51, MHDAnalysis.jl:94-174,237-255)
# --------------------------------------------------------------------------
def Curl(B1, B2, B3, grid: Grid):
    CT = grid.CT
    B1h, B2h, B3h = grid.rfft(B1), grid.rfft(B2), grid.rfft(B3)
    kx, ky, kz = grid.kr, grid.l, grid.m
    c1 = grid.irfft(CT(1j) * (ky * B3h - kz * B2h))
    c2 = grid.irfft(CT(1j) * (kz * B1h - kx * B3h))
    c3 = grid.irfft(CT(1j) * (kx * B2h - ky * B1h))
    return c1, c2, c3


def h_k(iv, jv, kv, grid: Grid):
    """MHDAnalysis.jl:94-101: pointwise (curl v).v * dV array (user sums)."""
    dV = (grid.Lx / grid.nx) * (grid.Ly / grid.ny) * (grid.Lz / grid.nz)
    c1, c2, c3 = Curl(iv, jv, kv, grid)
    return (c1 * iv + c2 * jv + c3 * kv) * grid.T(dV)


def VectorPotential(B1, B2, B3, grid: Grid):
    """MHDAnalysis.jl:129-174: Coulomb gauge A_k = i (k x B_k) / k^2."""
    CT = grid.CT
    B1h, B2h, B3h = grid.rfft(B1), grid.rfft(B2), grid.rfft(B3)
    kx, ky, kz, ik2 = grid.kr, grid.l, grid.m, grid.invKrsq
    A1 = grid.irfft(CT(1j) * (ky * B3h - kz * B2h) * ik2)
    A2 = grid.irfft(CT(1j) * (kz * B1h - kx * B3h) * ik2)
    A3 = grid.irfft(CT(1j) * (kx * B2h - ky * B1h) * ik2)
    return A1, A2, A3


def ScaleDecomposition(B1, B2, B3, grid: Grid, kf=(1, 5)):
    """MHDAnalysis.jl:54-82: rfft, keep k1 <= |k| <= k2 (|k| = sqrt(kr^2 + l^2 + m^2) in T), irfft; no dealias."""
    k1, k2 = min(kf), max(kf)
    kr = np.sqrt(grid.kr ** 2 + grid.l ** 2 + grid.m ** 2).astype(grid.T)
    K = ((k2 >= kr) & (kr >= k1)).astype(grid.T)
    return tuple(grid.irfft((grid.rfft(B) * K).astype(grid.CT)) for B in (B1, B2, B3))


def CF(V):
    """TurbStatTool.jl:67: CF(V) = fftshift(real(ifft(abs.(fft(V)).^2))) -- periodic autocorrelation of a cube."""
    return np.fft.fftshift(np.real(sfft.ifftn(np.abs(sfft.fftn(V.astype(np.float64))) ** 2))).astype(V.dtype)


def SFC(V):
    """TurbStatTool.jl:72: SFC(V) = 2*(mean(V) .- CF(V)), as written (the mean of V, not of V^2)."""
    return (2 * (np.mean(V) - CF(V))).astype(V.dtype)


def SF2_1D(Vx, Vz, Vy):
    """TurbStatTool.jl:90-120 (SF_2 1D), literally: 1-based loops, displacement measured from element (N/2, N/2, N/2), shells of
    round(|r|) (ties to even), 2R output bins (empty shells give 0/0 = NaN)."""
    nz, ny, nx = Vx.shape
    SFV = (SFC(Vx) + SFC(Vy) + SFC(Vz)).astype(np.float64)
    R = int(math.ceil(math.sqrt((nx // 2) ** 2 + (ny // 2) ** 2 + (nz // 2) ** 2)))
    mask, sfr = np.zeros(2 * R), np.zeros(2 * R)
    for k in range(1, nz + 1):
        for j in range(1, ny + 1):
            for i in range(1, nx + 1):
                kk = int(np.rint(math.sqrt((i - nx // 2) ** 2 + (j - ny // 2) ** 2 + (k - nz // 2) ** 2)))
                if kk > 0:
                    mask[kk - 1] += 1
                    sfr[kk - 1] += SFV[k - 1, j - 1, i - 1]
    with np.errstate(invalid="ignore", divide="ignore"):
        return sfr / mask


def h_m(ib, jb, kb, grid: Grid):
    """MHDAnalysis.jl:113-117: pointwise A.B (no dV)."""
    A1, A2, A3 = VectorPotential(ib, jb, kb, grid)
    return A1 * ib + A2 * jb + A3 * kb


def spectralline(A, grid: Grid):
    """MHDAnalysis.jl:237-255: Pk[r] += |A_k|^2, r = round(|k|)+1, HALF spectrum only,
    unnormalised rfft; kr[r] = r (bin index).  Returns (Pk, kr) 0-based arrays of
    length krmax = round(max|k| + 1)."""
    Ak = grid.rfft(A)
    kk = np.sqrt(grid.Krsq)
    krmax = int(np.rint(float(kk.max()) + 1))
    r = np.rint(kk).astype(np.int64)  # 0-based bin = round(|k|)
    Pk = np.zeros(krmax, dtype=np.float64)
    np.add.at(Pk, r.ravel(), (np.abs(Ak) ** 2).astype(np.float64).ravel())
    kr = np.zeros(krmax, dtype=grid.T)
    used = np.unique(r)
    kr[used] = used + 1
    return Pk.astype(grid.T), kr


# --------------------------------------------------------------------------
# Initial conditions
# --------------------------------------------------------------------------
def DivFreeSpectraMap(grid: Grid, theta, k_peak=0.0, P=1, k0=-5 / 3 / 2, b=1):
    """IC.jl:130-179 with the uniform random numbers injected: theta = rand(T, nkr, nl, nm)
    given as a NumPy array of shape (nm, nl, nkr) in [0, 1)."""
    T, CT = grid.T, grid.CT
    kx, ky, kz = grid.kr, grid.l, grid.m
    dx, dy, dz = grid.dx, grid.dy, grid.dz
    with np.errstate(divide="ignore", invalid="ignore"):
        kinv = np.sqrt(grid.invKrsq)
        k = np.sqrt(grid.Krsq)
        kperp = np.sqrt(kx ** 2 + ky ** 2) + 0 * kz
        dkm2 = 1 / (k + 1) ** 2
        Fk = k ** T(k0)
        Fk[0, 0, 0] = 0
        Fk[..., 0] = 0                                   # Fk[1,:,:] .= 0  (kr = 0 plane)
        Fk[k < k_peak] = 0
        intF = float(np.sum((Fk * dkm2).astype(np.float64)))
        Aamp = math.sqrt(P * 3 * (grid.Lx / dx) * (grid.Ly / dy) * (grid.Lz / dz) / intF * (1 / dx / dy / dz))
        Fk = (Fk * T(Aamp)).astype(T)
        e2x = kx * kz / kperp * kinv
        e2y = ky * kz / kperp * kinv
        e2z = -kperp * kinv
    e2x[np.isnan(e2x)] = 0
    e2y[np.isnan(e2y)] = 0
    eith = np.exp(1j * theta.astype(np.float64) * 2 * math.pi).astype(CT)
    Fxh = (Fk * eith * e2x).astype(CT)
    Fyh = (Fk * eith * e2y).astype(CT)
    Fzh = (Fk * eith * e2z).astype(CT)
    for F in (Fxh, Fyh, Fzh):
        grid.dealias(F)
    return grid.irfft(Fxh), grid.irfft(Fyh), grid.irfft(Fzh)


def taylor_green_ic(grid: Grid, with_b=True):
    """SURVEY 8d config 1/2 IC (ours; the reference ships no TG initial condition):
    u = (sin x cos y cos z, -cos x sin y cos z, 0),
    b = (cos x sin y sin z, sin x cos y sin z, -2 sin x sin y cos z)."""
    T = grid.T
    X = grid.x.astype(np.float64).reshape(1, 1, -1)
    Y = grid.y.astype(np.float64).reshape(1, -1, 1)
    Z = grid.z.astype(np.float64).reshape(-1, 1, 1)
    ux = (np.sin(X) * np.cos(Y) * np.cos(Z)).astype(T)
    uy = (-np.cos(X) * np.sin(Y) * np.cos(Z)).astype(T)
    uz = np.zeros_like(ux)
    if not with_b:
        return ux, uy, uz
    bx = (np.cos(X) * np.sin(Y) * np.sin(Z)).astype(T)
    by = (np.sin(X) * np.cos(Y) * np.sin(Z)).astype(T)
    bz = (-2 * np.sin(X) * np.sin(Y) * np.cos(Z)).astype(T)
    return ux, uy, uz, bx, by, bz


def random_phase_ic(grid: Grid, seed, k0=-5 / 6, P=1, k_peak=0.0):
    """DivFreeSpectraMap with phases from default_rng(seed), generated in
    the (nkr, nl, nm) column-major order == NumPy shape (nm, nl, nkr) C-order."""
    rng = np.random.default_rng(seed)
    theta = rng.random((grid.nm, grid.nl, grid.nkr), dtype=np.float64).astype(grid.T)
    return DivFreeSpectraMap(grid, theta, k_peak=k_peak, P=P, k0=k0)


# --------------------------------------------------------------------------
# helpers for parity tests
# --------------------------------------------------------------------------
def rel_l2(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    den = float(np.linalg.norm(b.astype(np.complex128).ravel()))
    num = float(np.linalg.norm((a.astype(np.complex128) - b.astype(np.complex128)).ravel()))
    return num / den if den > 0 else num
