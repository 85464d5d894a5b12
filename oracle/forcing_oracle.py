"""CPU oracle for the forcing modules and the divergence corrections -- TEST INFRASTRUCTURE ONLY.

Literal NumPy restatement of
  * A99ForceDriving! / SetUpFk / GetA99vars_And_function          (src/pgen/A99ForceDriving.jl:5-60, 93-127)
  * A99GPU.A99ForceDriving! / A99GPU.SetUpFk!                      (src/pgen/A99ForceDriving_GPU.jl:14-130)
  * VPSolver.DivBCorrection! / DivVCorrection!                     (src/Solver/VPSolver.jl:61-137)

PARITY UNPINNED like the rest of the oracle (the reference has no tests or fixtures and Julia cannot run here).  In
addition Julia's random streams (`Base.rand`, `CUDA.rand`) cannot be reproduced: the random draws are taken from a
counter-based Philox4x32-10 generator (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11;
pinned here by the Random123 known-answer vectors, tests/test_forcing_oracle.py) numbered exactly like the device
stream: counter = (index of the mode in the (nkr, nl, nm) array, forcing call, 0|1), key = seed; the reference's draws of
one call map onto the four words of that counter in the order the reference makes them.

Only tests/ may import this module.
Array convention as in mhdflows_oracle.py: Julia (nkr, nl, nm) == NumPy C-order (nm, nl, nkr).
"""
from __future__ import annotations

import math

import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32(ctr, key, rounds=10):
    """Philox4x32-R on arrays: ctr = 4 uint32 arrays (or scalars), key = (k0, k1) -> 4 uint32 arrays."""
    c = [np.asarray(x, dtype=np.uint64) & MASK for x in ctr]
    c = list(np.broadcast_arrays(*c))
    k0, k1 = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
    for _ in range(rounds):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [(p1 >> np.uint64(32)) ^ c[1] ^ np.uint64(k0), p1 & MASK, (p0 >> np.uint64(32)) ^ c[3] ^ np.uint64(k1), p0 & MASK]
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return [x.astype(np.uint32) for x in c]


class PhiloxField:
    """Four uniforms in [0,1) per mode and forcing call, numbered like the device stream (csrc/kernels.cuh::a99_uniforms):
    Float32: word j >> 8 scaled by 2^-24; Float64: words (2j, 2j+1) of the two blocks (sub = 0, 1), top 53 bits."""

    def __init__(self, seed, grid):
        self.seed = int(seed)
        self.T = grid.T
        nkr, nl, nm = grid.nkr, grid.nl, grid.nm
        ix = np.arange(nkr, dtype=np.uint64).reshape(1, 1, nkr)
        iy = np.arange(nl, dtype=np.uint64).reshape(1, nl, 1)
        iz = np.arange(nm, dtype=np.uint64).reshape(nm, 1, 1)
        self.mode = ix + np.uint64(nkr) * (iy + np.uint64(nl) * iz)

    def uniforms(self, call):
        key = (self.seed & 0xFFFFFFFF, self.seed >> 32)
        lo, hi = self.mode & MASK, self.mode >> np.uint64(32)
        c2, c3 = call & 0xFFFFFFFF, ((call >> 32) << 1) & 0xFFFFFFFF
        a = philox4x32((lo, hi, c2, c3), key)
        if self.T is np.float32:
            return [((w >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)) for w in a]
        b = philox4x32((lo, hi, c2, c3 | 1), key)
        w = a + b
        out = []
        for j in range(4):
            x = (w[2 * j].astype(np.uint64) << np.uint64(32)) | w[2 * j + 1].astype(np.uint64)
            out.append((x >> np.uint64(11)).astype(np.float64) * 2.0 ** -53)
        return out


class NDVars:
    """ND_vars (NegativeDamping.jl:8-13): P and the real profiles fx, fy, fz."""

    def __init__(self, grid):
        z = lambda: np.zeros((grid.nz, grid.ny, grid.nx), dtype=grid.T)
        self.P, self.fx, self.fy, self.fz = 0.0, z(), z(), z()


def SetUpND(prob, P, fx, fy, fz):
    """SetUpND!(prob, P, fx, fy, fz) (NegativeDamping.jl:14-21)."""
    uv = prob.vars.usr_vars
    uv.P = P
    uv.fx[...], uv.fy[...], uv.fz[...] = fx, fy, fz


def NDForceDriving(N, sol, t, clock, vars, params, grid):
    """NDForceDriving! (NegativeDamping.jl:23-45): N_ui += A rfft(f_i u_i), A = P / ((sum|ux^2 fx| + sum|uy^2 fy| + sum|uz^2 fz|) dx dy dz)
    with the vars.u* the advection of this evaluation just refreshed."""
    uv = vars.usr_vars
    u = (vars.ux, vars.uy, vars.uz)
    f = (uv.fx, uv.fy, uv.fz)
    integral = sum(float(np.sum(np.abs(ui ** 2 * fi))) for ui, fi in zip(u, f)) * grid.dx * grid.dy * grid.dz
    A = uv.P / integral
    for fi, ui, ind in zip(f, u, (params.ux_ind, params.uy_ind, params.uz_ind)):
        vars.nonlinh1[...] = 0
        vars.nonlin1[...] = fi * ui
        vars.nonlinh1[...] = grid.rfft(vars.nonlin1)
        N[ind] += (A * vars.nonlinh1).astype(grid.CT)


class A99Vars:
    """A99_vars{Atrans,T} (A99ForceDriving.jl:5-16): A, b scalars of type T; Fk, e1x, e1y, e2x, e2y, e2z, gi, e^{iθ}
    Complex{T} arrays (nkr, nl, nm)."""

    def __init__(self, grid):
        T = grid.T
        self.A, self.b = T(1.0), T(1.0)
        z = lambda: np.zeros((grid.nm, grid.nl, grid.nkr), dtype=grid.CT)
        self.Fk, self.e1x, self.e1y, self.e2x, self.e2y, self.e2z, self.gi, self.eith = (z() for _ in range(8))
        self.rng = None
        self.calls = 0


def SetUpFk(prob, kf=2, P=1, sigma2=1):
    """SetUpFk(prob; kf, P, σ²) (A99ForceDriving.jl:93-127).  `k`, `k⁻¹` come from the T-typed Krsq / invKrsq arrays; with the
    default Int kf, σ² everything stays in T."""
    g = prob.grid
    T = g.T
    uv = prob.vars.usr_vars
    kx, ky, kz = g.kr, g.l, g.m
    dx, dy, dz = g.dx, g.dy, g.dz
    with np.errstate(divide="ignore", invalid="ignore"):
        kinv = np.sqrt(g.invKrsq)
        k = np.sqrt(g.Krsq)
        kperp = np.sqrt(kx ** 2 + ky ** 2)              # shape (1, nl, nkr)
        dkm2 = 1 / (k + T(1)) ** 2
        kf_, s2_ = (T(kf) if float(kf).is_integer() else kf), (T(sigma2) if float(sigma2).is_integer() else sigma2)
        integral = np.sum(np.exp(-(k - kf_) ** 2 / s2_) * dkm2)
        A = math.sqrt(P * 3 * (g.Lx / dx) * (g.Ly / dy) * (g.Lz / dz) / float(integral) * (1 / dx / dy / dz))
        Fk = A * np.sqrt(np.exp(-(k - kf_) ** 2 / s2_) / 2 / math.pi) * kinv
        Fk[:, :, 0] = 0                                  # Fk[1,:,:] .= 0  (FourierFlows issue 326)
        e1x = ky / kperp                                 # (nkr, nl, 1) arrays in the reference
        e1y = -kx / kperp
        e2x = kx * kz / kperp * kinv
        e2y = ky * kz / kperp * kinv
        e2z = -kperp * kinv
    for a in (e1x, e1y, e2x, e2y):
        a[np.isnan(a)] = 0
    uv.Fk[...] = Fk
    # copyto!(usr_vars.e1x, e1x) with a (nkr, nl, 1) source fills the first nkr*nl elements = the first z plane only
    uv.e1x[...], uv.e1y[...] = 0, 0
    uv.e1x[0], uv.e1y[0] = e1x[0], e1y[0]
    uv.e2x[...], uv.e2y[...], uv.e2z[...] = e2x, e2y, e2z
    return A


def A99ForceDriving(N, sol, t, clock, vars, params, grid):
    """A99ForceDriving! (A99ForceDriving.jl:33-60).  Φ is the COMPLEX scratch array vars.nonlinh1: rand!(Φ) fills real and
    imaginary parts, so gi = -tanh(b(Φ-π/2))/tanh(bπ/2) and √(1-gi²) are complex (restated as written)."""
    uv = vars.usr_vars
    T, CT = grid.T, grid.CT
    A, b = uv.A, uv.b
    r = uv.rng.uniforms(uv.calls)
    uv.calls += 1
    two_pi, pi = T(2 * math.pi), T(math.pi)
    uv.eith[...] = np.exp(1j * (r[0] * two_pi)).astype(CT)
    Phi = vars.nonlinh1
    Phi[...] = (r[1] + 1j * r[2]).astype(CT)            # rand!(Φ)
    Phi *= pi
    uv.gi[...] = (-np.tanh(b * (Phi - pi / 2)) / np.tanh(b * pi / 2)).astype(CT)
    N[params.ux_ind] += A * uv.Fk * uv.eith * uv.gi * uv.e1x
    N[params.uy_ind] += A * uv.Fk * uv.eith * uv.gi * uv.e1y
    uv.eith[...] = np.exp(1j * (r[3] * two_pi)).astype(CT)
    uv.gi[...] = np.sqrt(1 - uv.gi ** 2).astype(CT)
    N[params.ux_ind] += A * uv.Fk * uv.eith * uv.gi * uv.e2x
    N[params.uy_ind] += A * uv.Fk * uv.eith * uv.gi * uv.e2y
    N[params.uz_ind] += A * uv.Fk * uv.eith * uv.gi * uv.e2z


class A99GPUVars:
    """A99GPU.A99_vars{T} (A99ForceDriving_GPU.jl:7-12)."""

    def __init__(self, grid):
        T = grid.T
        self.A, self.b, self.sigma2, self.kf = T(1.0), T(1.0), T(1.0), T(1.0)
        self.rng = None
        self.calls = 0


def SetUpFk_GPU(prob, kf=2.0, P=1.0, sigma=1.0, b=1.0):
    """A99GPU.SetUpFk! (A99ForceDriving_GPU.jl:25-48); note `usr_vars.kf = T(b)` (:45)."""
    g = prob.grid
    T = g.T
    k = np.sqrt(g.Krsq).astype(np.float64)               # Float32 k against Float64 kf, σ -> Float64 terms
    dkm2 = 1 / (k + 1) ** 2
    integral = np.sum(np.exp(-(k - kf) ** 2 / sigma ** 2) * dkm2)
    A = math.sqrt(P * 3 * (g.Lx / g.dx) * (g.Ly / g.dy) * (g.Lz / g.dz) / integral * (1 / g.dx / g.dy / g.dz))
    uv = prob.vars.usr_vars
    uv.A, uv.sigma2, uv.b, uv.kf = T(A), T(sigma ** 2), T(b), T(b)


def A99ForceDriving_GPU(N, sol, t, clock, vars, params, grid):
    """A99GPU.A99ForceDriving! + A99Force_Driving_CUDA! (A99ForceDriving_GPU.jl:49-130), one array expression per kernel
    statement.  The three scalar rand() calls of a thread map onto words 0, 1, 3 of that mode's counter."""
    uv = vars.usr_vars
    T = grid.T
    A, b, kf, s2 = uv.A, uv.b, uv.kf, uv.sigma2
    r = uv.rng.uniforms(uv.calls)
    uv.calls += 1
    kx, ky, kz = grid.kr, grid.l, grid.m
    pi = T(math.pi)
    with np.errstate(divide="ignore", invalid="ignore"):
        k = np.sqrt(kx ** 2 + ky ** 2 + kz ** 2)
        kperp = np.sqrt(kx ** 2 + kz ** 2) + 0 * k
        kinv = np.where(k > 0, 1 / k, 0).astype(T)
        Fk = A * np.sqrt(np.exp(-(k - kf) ** 2 / s2) / 2 / pi) * kinv
        ok = kperp > 0
        e1x = np.where(ok, kz / kperp, 0)
        e1z = np.where(ok, -kx / kperp, 0)
        e2x = np.where(ok, kx * ky / kperp * kinv, 0)
        e2y = -kperp * kinv
        e2z = np.where(ok, kz * ky / kperp * kinv, 0)
    eith = np.exp(1j * (r[0] * 2 * pi))
    Phi = r[1] * pi
    gi = -np.tanh(b * (Phi - pi / 2)) / np.tanh(b * pi / 2)
    gi = np.where(np.abs(gi) >= 1, np.sign(gi), gi)
    N[params.ux_ind] += (A * Fk * eith * gi * e1x).astype(grid.CT)
    N[params.uz_ind] += (A * Fk * eith * gi * e1z).astype(grid.CT)
    eith = np.exp(1j * (r[3] * 2 * pi))
    gj = np.sqrt(1 - gi ** 2)
    N[params.ux_ind] += (A * Fk * eith * gj * e2x).astype(grid.CT)
    N[params.uy_ind] += (A * Fk * eith * gj * e2y).astype(grid.CT)
    N[params.uz_ind] += (A * Fk * eith * gj * e2z).astype(grid.CT)
    for ind in (params.ux_ind, params.uy_ind, params.uz_ind):      # x == 1 || x == nx: keep the real part only
        for x in (0, grid.nkr - 1):
            N[ind][:, :, x] = N[ind][:, :, x].real


def _div_correction(prob, inds, names):
    g, v = prob.grid, prob.vars
    ki, kj, kk = g.kr, g.l, g.m
    xh, yh, zh = (prob.sol[i] for i in inds)
    phi = (-1j * (ki * xh + kj * yh + kk * zh)).astype(g.CT)
    phi = (phi * g.invKrsq).astype(g.CT)
    xh -= 1j * ki * phi
    yh -= 1j * kj * phi
    zh -= 1j * kk * phi
    for name, fh in zip(names, (xh, yh, zh)):
        getattr(v, name)[...] = g.irfft(fh.copy())


def DivBCorrection(prob):
    """VPSolver.DivBCorrection! (VPSolver.jl:61-99)."""
    p = prob.params
    _div_correction(prob, (p.bx_ind, p.by_ind, p.bz_ind), ("bx", "by", "bz"))


def DivVCorrection(prob):
    """VPSolver.DivVCorrection! (VPSolver.jl:101-137)."""
    p = prob.params
    _div_correction(prob, (p.ux_ind, p.uy_ind, p.uz_ind), ("ux", "uy", "uz"))
