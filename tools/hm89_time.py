"""HM89TimeStepper (Problem(...; EMHD = true, stepper = "HM89")) on one GPU: ms per step and fixed-point iterations per step next to
RK4 on the same EMHD problem.  The HM89 step costs 3 + iterations Hall-term evaluations; its point is the time step (the explicit
steppers need the EMHD CFL dt ~ dx^2, integrator.jl:186-193).  usage: hm89_time.py [n] [steps]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mhdflows_jl_b200 as M  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
for stepper in ("RK4", "HM89"):
    p = M.Problem(M.GPU(), nx=n, B_field=True, EMHD=True, stepper=stepper, eta=1e-4, dt=0.0)
    M.SetUpRandomPhaseIC(p, seed_b=7, P=1e-2, k_peak=0.0)
    p.calcN()                                   # vars.curlB for the first getCFL!
    dt_cfl = M.getCFL(p, 1e9, Coef=0.3)
    p.clock.dt = dt_cfl * (8.0 if stepper == "HM89" else 1.0)
    M.stepforward(p, 1)
    iters = []
    t0 = time.perf_counter()
    for _ in range(steps):
        M.stepforward(p, 1)
        iters.append(p.stepper_stats()[0])
    ms = (time.perf_counter() - t0) / steps * 1e3
    KE, ME = p.energy(M.FRESH)
    print(f"emhd {n}^3 {stepper}: dt = {p.clock.dt:.3e} ({p.clock.dt / dt_cfl:.0f} x CFL dt)  {ms:.2f} ms/step  fixed-point iterations {iters}  ME {ME:.6e}")
    p.close()
