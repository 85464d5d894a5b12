"""GPU stand-in for the reference's CUDA.jl path ("v0 composition") -- MEASUREMENT BASELINE ONLY.

The reference's GPU mode cannot be installed here (Julia, CUDA.jl and FourierFlows.jl are absent and there is no network).
This module restates what that path executes on the device: the literal op sequence of `MHDcalcN!` / `HDcalcN!`
(src/pgen.jl:153-181; src/Solver/MHDSolver.jl:27-177, 330-351; src/Solver/HDSolver.jl:25-108) -- 6 (3) c2r + 30 (21) r2c
cuFFT transforms per RHS evaluation with `deepcopy` before every c2r, scratch zero-fills, one unfused broadcast kernel per
`@.` line -- and FourierFlows' RK4TimeStepper (sol1, RHS1..4; mirror src/DyeModule.jl:62-91), in PyTorch eager mode on CUDA:
torch.fft = cuFFT, every broadcast = its own elementwise kernel(s), exactly the structure CUDA.jl generates.

None of the product's kernels is used and the product never imports this file; only bench.py's `gpu_baseline` leg does.
It is checked against the CPU oracle in tests/test_gpu_baseline.py (so the baseline computes the same thing).
"""
from __future__ import annotations

import math

import torch


class TorchV0:
    LSRK_A = (0.0, -567301805773.0 / 1357537059087.0, -2404267990393.0 / 2016746695238.0, -3550918686646.0 / 2091501179385.0,
              -1275806237668.0 / 842570457699.0)
    LSRK_B = (1432997174477.0 / 9575080441755.0, 5161836677717.0 / 13612068292357.0, 1720146321549.0 / 2090206949498.0,
              3134564353537.0 / 4481467310338.0, 2277821191437.0 / 14882151754819.0)

    def __init__(self, n, kind="mhd", nu=0.0, eta=0.0, dt=0.0, device="cuda", L=2 * math.pi, stepper="RK4"):
        self.n, self.kind, self.nu, self.eta, self.dt, self.stepper = n, kind, float(nu), float(eta), float(dt), stepper
        self.dev = torch.device(device)
        f32, dev = torch.float32, self.dev
        self.nkr = n // 2 + 1
        k1 = torch.arange(self.nkr, dtype=torch.float64) * (2 * math.pi / L)
        kf = torch.fft.fftfreq(n, 1.0 / n).to(torch.float64) * (2 * math.pi / L)
        self.kr = k1.to(f32).reshape(1, 1, -1).to(dev)
        self.l = kf.to(f32).reshape(1, -1, 1).to(dev)
        self.m = kf.to(f32).reshape(-1, 1, 1).to(dev)
        self.ks = (self.kr, self.l, self.m)
        self.Krsq = self.kr ** 2 + self.l ** 2 + self.m ** 2          # 3D arrays like grid.Krsq / grid.invKrsq
        self.invKrsq = 1 / self.Krsq
        self.invKrsq[0, 0, 0] = 0
        self.Krsq64 = self.Krsq.to(torch.float64)
        iL = math.floor((1 - 1 / 3) / 2 * n) + 1
        iR = math.ceil((1 + 1 / 3) / 2 * n)
        self.kralias, self.lalias = slice(iL - 1, self.nkr), slice(iL - 1, iR)
        self.Nl = 6 if kind == "mhd" else 3
        z = lambda: torch.zeros((n, n, n), dtype=f32, device=dev)
        zc = lambda *lead: torch.zeros(lead + (n, n, self.nkr), dtype=torch.complex64, device=dev)
        self.vars = [z() for _ in range(self.Nl)]                      # ux,uy,uz[,bx,by,bz]
        self.nonlin1, self.nonlinh1 = z(), zc()
        self.sol = zc(self.Nl)
        if stepper == "RK4":
            self.sol1 = zc(self.Nl)
            self.RHS = [zc(self.Nl) for _ in range(4)]
        else:                                                          # FourierFlows LSRK54TimeStepper: S2 and RHS
            self.S2 = zc(self.Nl)
            self.RHS = [zc(self.Nl)]
        self.t, self.step = 0.0, 0

    # mul!(yh, rfftplan, y) / ldiv!(y, rfftplan, deepcopy(yh))
    def rfft(self, f):
        return torch.fft.rfftn(f, dim=(0, 1, 2))

    def irfft(self, fh):
        return torch.fft.irfftn(fh.clone(), s=(self.n, self.n, self.n), dim=(0, 1, 2))

    def dealias(self, fh):                                             # three strided fills (FourierFlows.dealias!)
        fh[..., self.kralias] = 0
        fh[..., self.lalias, :] = 0
        fh[..., self.lalias, :, :] = 0

    def set_ic(self, fields):
        for i, f in enumerate(fields[: self.Nl]):
            self.vars[i].copy_(torch.as_tensor(f).to(self.dev))
            self.sol[i] = self.rfft(self.vars[i])

    def _ui_update(self, N, a):                                        # MHDSolver.jl:27-103 / HDSolver.jl:25-93
        ks, us = self.ks, self.vars[:3]
        bs = self.vars[3:6] if self.kind == "mhd" else None
        ka, kinv2 = ks[a], self.invKrsq
        dudt = N[a]
        dudt.mul_(0)
        for i in range(3):
            for j in range(i, 3):
                self.nonlin1.mul_(0)
                self.nonlinh1.mul_(0)
                if bs is not None:
                    self.nonlin1.copy_(bs[i] * bs[j] - us[i] * us[j])
                    sgn = 1j
                else:
                    self.nonlin1.copy_(us[i] * us[j])
                    sgn = -1j
                self.nonlinh1.copy_(self.rfft(self.nonlin1))
                dudt.add_((sgn * ks[i] * ((1.0 if a == j else 0.0) - ka * ks[j] * kinv2)) * self.nonlinh1)
                if i != j:
                    dudt.add_((sgn * ks[j] * ((1.0 if a == i else 0.0) - ka * ks[i] * kinv2)) * self.nonlinh1)
        self.nonlinh1.copy_(self.rfft(us[a]))
        dudt.add_((-self.Krsq64 * self.nu * self.nonlinh1).to(torch.complex64))

    def _bi_update(self, N, a):                                        # MHDSolver.jl:106-177
        ks, us, bs = self.ks, self.vars[:3], self.vars[3:6]
        dbdt = N[3 + a]
        dbdt.mul_(0)
        for j in range(3):
            if a != j:
                self.nonlin1.copy_(us[a] * bs[j] - bs[a] * us[j])
                self.nonlinh1.copy_(self.rfft(self.nonlin1))
                dbdt.add_((1j * ks[j]) * self.nonlinh1)
        self.nonlinh1.copy_(self.rfft(bs[a]))
        dbdt.add_((-self.Krsq64 * self.eta * self.nonlinh1).to(torch.complex64))

    def calcN(self, N, sol):                                           # pgen.jl:153-162 / 173-181
        self.dealias(sol)
        for i in range(self.Nl):
            self.vars[i].copy_(self.irfft(sol[i]))
        for a in range(3):
            self._ui_update(N, a)
        if self.kind == "mhd":
            for a in range(3):
                self._bi_update(N, a)

    def stepforward(self):                                             # FourierFlows RK4 (mirror DyeModule.jl:62-91)
        if self.stepper != "RK4":                                      # LSRK54 (Carpenter & Kennedy 1994)
            self.S2.mul_(0)
            for a, b in zip(self.LSRK_A, self.LSRK_B):
                self.calcN(self.RHS[0], self.sol)
                self.S2.copy_(a * self.S2 + self.dt * self.RHS[0])
                self.sol.add_(b * self.S2)
            self.t += self.dt
            self.step += 1
            return
        dt, sol, R = self.dt, self.sol, self.RHS
        self.calcN(R[0], sol)
        self.sol1.copy_(sol + (dt / 2) * R[0])
        self.calcN(R[1], self.sol1)
        self.sol1.copy_(sol + (dt / 2) * R[1])
        self.calcN(R[2], self.sol1)
        self.sol1.copy_(sol + dt * R[2])
        self.calcN(R[3], self.sol1)
        sol.add_(dt * (R[0] / 6 + R[1] / 3 + R[2] / 3 + R[3] / 6))
        self.t += dt
        self.step += 1
