"""NumPy model of the PRODUCT's formulation -- TEST INFRASTRUCTURE ONLY.

The CUDA library does not follow the reference's 36/24/51-FFT op sequence; it uses the
minimum-FFT fused form on a compact (dealiased-modes-only) state:

  MHD  (6 c2r + 9 r2c):  T_ij = b_i b_j - u_i u_j,  D_j = sum_i i k_i T^_ij,
        N_a     = D_a - k_a (k.D)/k^2 - nu k^2 u^sym_a  [- nu k^(2 n_nu) u^sym_a if n_nu > 1]
        E = u x b,  N_{3+a} = i (k x E^)_a - eta k^2 b^sym_a
  HD   (3 c2r + 6 r2c):  T_ij = -u_i u_j only
  EMHD (21 c2r + 3 r2c): A = curl B, N_i = F[ sum_j A_j d_j B_i - b^stale_j d_j A_i ]

where ^sym is the kr=0-plane Hermitian symmetrisation that the reference's
`rfft(irfft(sol))` diffusion operand implies.  This module states that
formulation in NumPy so tests can prove it equal to the literal oracle
(oracle/mhdflows_oracle.py) to round-off before the same formulas are trusted on the GPU.
Never imported by the product path.
"""
from __future__ import annotations

import numpy as np

from . import mhdflows_oracle as O


def sym_kr0(fh, grid: O.Grid):
    """rfft(irfft(fh)) for a dealiased fh: Hermitian-symmetrise the kr=0 plane."""
    out = fh.copy()
    p = fh[..., 0]
    mir = np.conj(np.roll(np.roll(p[..., ::-1, ::-1], 1, axis=-1), 1, axis=-2))
    out[..., 0] = 0.5 * (p + mir)
    return out


def rhs_mhd(sol, grid: O.Grid, nu, eta, n_nu=0, hd=False):
    """One RHS evaluation on a (masked) copy of sol; returns N masked to the retained band."""
    CT = grid.CT
    s = grid.dealias(sol.copy())
    kx, ky, kz, ik2, k2 = grid.kr, grid.l, grid.m, grid.invKrsq, grid.Krsq
    ks = (kx, ky, kz)
    u = [grid.irfft(s[i].copy()) for i in range(3)]
    b = None if hd else [grid.irfft(s[3 + i].copy()) for i in range(3)]
    N = np.zeros_like(s)
    D = [np.zeros_like(s[0]) for _ in range(3)]
    for i in range(3):
        for j in range(i, 3):
            T = -(u[i] * u[j]) if hd else (b[i] * b[j] - u[i] * u[j])
            Th = grid.rfft(T)
            D[j] += CT(1j) * ks[i] * Th
            if i != j:
                D[i] += CT(1j) * ks[j] * Th
    kD = kx * D[0] + ky * D[1] + kz * D[2]
    k2d = k2.astype(np.float64)
    for a in range(3):
        us = sym_kr0(s[a], grid)
        N[a] = D[a] - ks[a] * kD * ik2 + (-k2d * nu * us).astype(CT)
        if n_nu > 1:
            N[a] += (-(k2d ** n_nu) * nu * us).astype(CT)
    if not hd:
        E = (u[1] * b[2] - u[2] * b[1], u[2] * b[0] - u[0] * b[2], u[0] * b[1] - u[1] * b[0])
        Eh = [grid.rfft(e) for e in E]
        C = (ky * Eh[2] - kz * Eh[1], kz * Eh[0] - kx * Eh[2], kx * Eh[1] - ky * Eh[0])
        for a in range(3):
            bs = sym_kr0(s[3 + a], grid)
            N[3 + a] = CT(1j) * C[a] + (-k2d * eta * bs).astype(CT)
    return grid.dealias(N), u, b


def rhs_emhd(sol, grid: O.Grid, b_stale):
    """EMHD gradient form; returns (N masked, fresh real b, A)."""
    CT = grid.CT
    s = grid.dealias(sol.copy())
    ks = (grid.kr, grid.l, grid.m)
    Bh = [s[0], s[1], s[2]]
    Ah = [CT(1j) * (ks[1] * Bh[2] - ks[2] * Bh[1]),
          CT(1j) * (ks[2] * Bh[0] - ks[0] * Bh[2]),
          CT(1j) * (ks[0] * Bh[1] - ks[1] * Bh[0])]
    A = [grid.irfft(a.copy()) for a in Ah]
    N = np.zeros_like(s)
    for i in range(3):
        acc = np.zeros_like(A[0])
        for j in range(3):
            dB = grid.irfft(CT(1j) * ks[j] * Bh[i])
            dA = grid.irfft(CT(1j) * ks[j] * Ah[i])
            acc += A[j] * dB - b_stale[j] * dA
        N[i] = grid.rfft(acc)
    b_new = [grid.irfft(x.copy()) for x in Bh]
    return grid.dealias(N), b_new, A


def rhs_emhd_div(sol, grid: O.Grid, b_stale, bh_stale):
    """EMHD divergence form (SURVEY A.6 iii): 7 c2r + 12 r2c instead of 24 + 3.
        N_i = sum_j i k_j F[A_j B_i - b^st_j A_i] + F[A_i div(b^st)]
    `b_stale` is the stale real-space b (vars.b*), `bh_stale` its spectrum (the previous stage input), needed for
    div(b^st).  Groundwork for the round-2 CUDA path.  Exact away from the truncation edge; ~1e-10 per step off on
    broadband fields (unpaired edge modes alias differently) -- see tests/test_fused_model.py."""
    CT = grid.CT
    s = grid.dealias(sol.copy())
    ks = (grid.kr, grid.l, grid.m)
    Bh = [s[0], s[1], s[2]]
    Ah = [CT(1j) * (ks[1] * Bh[2] - ks[2] * Bh[1]),
          CT(1j) * (ks[2] * Bh[0] - ks[0] * Bh[2]),
          CT(1j) * (ks[0] * Bh[1] - ks[1] * Bh[0])]
    A = [grid.irfft(a.copy()) for a in Ah]                     # 3 c2r
    B = [grid.irfft(x.copy()) for x in Bh]                     # 3 c2r (also the next stale b)
    divb = grid.irfft(CT(1j) * (ks[0] * bh_stale[0] + ks[1] * bh_stale[1] + ks[2] * bh_stale[2]))   # 1 c2r
    N = np.zeros_like(s)
    for i in range(3):
        acc = grid.rfft(A[i] * divb)                           # 3 r2c
        for j in range(3):
            acc = acc + CT(1j) * ks[j] * grid.rfft(A[j] * B[i] - b_stale[j] * A[i])   # 9 r2c
        N[i] = acc
    return grid.dealias(N), B, A, [x.copy() for x in Bh]


class FusedProblem:
    """3-register RK4 / 2N LSRK54 on the masked state, mirroring the CUDA library's stepping."""

    def __init__(self, oracle_prob: O.Problem):
        p = oracle_prob
        self.grid, self.flag, self.params = p.grid, p.flag, p.params
        self.stepper = p.stepper
        self.dt = p.clock.dt
        self.sol = p.grid.dealias(p.sol.copy())
        # stale real-space b for EMHD = the (undealiased) IC real field (IC.jl:86-90)
        self.b_stale = [p.vars.bx.copy(), p.vars.by.copy(), p.vars.bz.copy()] if p.flag.e else None
        self.bh_stale = [self.sol[i].copy() for i in range(3)] if p.flag.e else None   # spectrum of the stale b
        self.emhd_div_form = False
        self.last_real = None

    def rhs(self, s):
        g, pr = self.grid, self.params
        if self.flag.e:
            if self.emhd_div_form:
                N, bnew, A, bh = rhs_emhd_div(s, g, self.b_stale, self.bh_stale)
                self.bh_stale = bh
            else:
                N, bnew, A = rhs_emhd(s, g, self.b_stale)
            self.b_stale = bnew
            self.last_real = (A, bnew)
            return N
        N, u, b = rhs_mhd(s, g, pr.nu, pr.eta, pr.n_nu, hd=not self.flag.b)
        self.last_real = (u, b)
        return N

    def step(self):
        T = self.grid.T
        dt = T(self.dt)
        Y = self.sol
        if self.stepper == "RK4":
            k = self.rhs(Y)
            A = Y + (dt / T(6)) * k
            S = Y + (dt / T(2)) * k
            k = self.rhs(S)
            A += (dt / T(3)) * k
            S = Y + (dt / T(2)) * k
            k = self.rhs(S)
            A += (dt / T(3)) * k
            S = Y + dt * k
            k = self.rhs(S)
            self.sol = A + (dt / T(6)) * k
        else:
            S2 = np.zeros_like(Y)
            for i in range(5):
                k = self.rhs(Y)
                S2 = T(float(O.LSRK54_A[i])) * S2 + dt * k
                Y = Y + T(float(O.LSRK54_B[i])) * S2
            self.sol = Y
