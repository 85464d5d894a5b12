"""First-contact GPU diagnostics: prints errors of every layer against NumPy / the oracle without asserting,
so one gpurun call yields the whole picture."""
import sys, os, time, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import scipy.fft as sfft
import mhdflows_jl_b200 as M
from oracle import mhdflows_oracle as O


def rel(a, b):
    return O.rel_l2(a, b)


def fft_checks(nx, ny, nz, T=np.float32):
    p = M.Problem(M.GPU(), nx=nx, ny=ny, nz=nz, T=T, B_field=True)
    g = p.grid
    rng = np.random.default_rng(1)
    x = rng.standard_normal((nz, ny, nx)).astype(T)
    p.set_real(0, x)
    got = p.get_spectral(0)
    ref = sfft.rfftn(x.astype(np.float64), axes=(0, 1, 2))
    msk = g.retained_mask()
    ref[~msk] = 0
    e1 = rel(got, ref)
    back = p.get_real(0)
    refb = sfft.irfftn(ref, s=(nz, ny, nx), axes=(0, 1, 2))
    e2 = rel(back, refb)
    # spectral set/get roundtrip with non-Hermitian garbage
    s = (rng.standard_normal((nz, ny, nx // 2 + 1)) + 1j * rng.standard_normal((nz, ny, nx // 2 + 1))).astype(p.CT)
    p.set_spectral(1, s)
    g1 = p.get_spectral(1)
    sm = s.copy(); sm[~msk] = 0
    e3 = rel(g1, sm)
    r1 = p.get_real(1)
    e4 = rel(r1, sfft.irfftn(sm.astype(np.complex128), s=(nz, ny, nx), axes=(0, 1, 2)))
    print(f"fft {nx}x{ny}x{nz} {np.dtype(T).name}: r2c {e1:.2e}  c2r {e2:.2e}  pack {e3:.2e}  c2r(nonherm) {e4:.2e}  info={p.info()}")
    p.close()


def setup_pair(kind, n, T, stepper="RK4", turb=False, dt=None, nxyz=None):
    nx, ny, nz = nxyz if nxyz else (n, n, n)
    kw = dict(nx=nx, ny=ny, nz=nz, T=T, stepper=stepper)
    if kind == "mhd":
        kw.update(nu=2e-2, eta=3e-2, B_field=True, dt=dt or 4e-3)
    elif kind == "hd":
        kw.update(nu=2e-2, dt=dt or 4e-3)
    else:
        kw.update(B_field=True, EMHD=True, dt=dt or 2e-4)
    op = O.Problem(**kw)
    gp = M.Problem(M.GPU(), **kw)
    g = op.grid
    if turb:
        u = O.random_phase_ic(g, 1234); b = O.random_phase_ic(g, 5678)
    else:
        ic = O.taylor_green_ic(g); u, b = ic[:3], ic[3:]
    if kind == "emhd":
        O.SetUpProblemIC(op, bx=b[0], by=b[1], bz=b[2]); M.SetUpProblemIC(gp, bx=b[0], by=b[1], bz=b[2])
    elif kind == "mhd":
        O.SetUpProblemIC(op, *u, bx=b[0], by=b[1], bz=b[2]); M.SetUpProblemIC(gp, ux=u[0], uy=u[1], uz=u[2], bx=b[0], by=b[1], bz=b[2])
    else:
        O.SetUpProblemIC(op, *u); M.SetUpProblemIC(gp, ux=u[0], uy=u[1], uz=u[2])
    return op, gp


def rhs_check(kind, n, T=np.float32, turb=True, nxyz=None):
    op, gp = setup_pair(kind, n, T, turb=turb, nxyz=nxyz)
    N = np.zeros_like(op.sol)
    op.calcN(N, op.sol, 0.0, op.clock, op.vars, op.params, op.grid)
    op.grid.dealias(N)
    got = gp.calcN()
    errs = [rel(got[i], N[i]) for i in range(op.Nl)]
    print(f"calcN {kind} {nxyz or n} {np.dtype(T).name} turb={turb}: per-field rel L2 = " + " ".join(f"{e:.2e}" for e in errs))
    gp.close()


def step_check(kind, n, T=np.float32, stepper="RK4", nsteps=10, turb=True):
    op, gp = setup_pair(kind, n, T, stepper=stepper, turb=turb)
    out = []
    for s in range(nsteps):
        O.stepforward(op)
        M.stepforward(gp)
        if s in (0, nsteps // 2, nsteps - 1):
            ref = op.grid.dealias(op.sol.copy())
            out.append(f"step{s+1}:{rel(gp.sol, ref):.2e}")
    ko = O.ProbDiagnostic(op, rounded=False)
    kg = gp.energy(M.STALE)
    kf = gp.energy(M.FRESH)
    print(f"steps {kind} {n} {np.dtype(T).name} {stepper}: " + " ".join(out) + f" | oracle E={ko} stale={kg} fresh={kf} t={gp.clock.t:.5f}/{op.clock.t:.5f}")
    gp.close()


def timing(kind, n, stepper="RK4", nsteps=5):
    kw = dict(nx=n, nu=1e-3, eta=1e-3, dt=1e-3, stepper=stepper)
    if kind == "mhd": kw.update(B_field=True)
    if kind == "emhd": kw.update(B_field=True, EMHD=True)
    gp = M.Problem(M.GPU(), **kw)
    g = gp.grid
    ic = O.taylor_green_ic(g)
    if kind == "emhd":
        M.SetUpProblemIC(gp, bx=ic[3], by=ic[4], bz=ic[5])
    elif kind == "mhd":
        M.SetUpProblemIC(gp, ux=ic[0], uy=ic[1], uz=ic[2], bx=ic[3], by=ic[4], bz=ic[5])
    else:
        M.SetUpProblemIC(gp, ux=ic[0], uy=ic[1], uz=ic[2])
    gp.step_timed(2)
    ms = gp.step_timed(nsteps) / nsteps
    gp.profile(True)
    gp.step_timed(nsteps)
    pr = gp.profile_get()
    gp.profile(False)
    tot = sum(v[0] for v in pr.values())
    S = 8 * (n // 2 + 1) * n * n
    alg = {"mhd": 384, "hd": 216, "emhd": 424}[kind] * S if stepper == "RK4" else 450 * S
    print(f"time {kind} {n}^3 {stepper}: {ms:.3f} ms/step  {n**3/ms*1e3:.3e} pts*steps/s  contract-roofline frac={alg/(ms*1e-3)/6555.2e9:.3f}  mem={gp.info()['bytes_device']/2**30:.2f} GiB")
    print("   per class ms/step: " + "  ".join(f"{k}={v[0]/nsteps:.3f}({v[1]//nsteps})" for k, v in pr.items() if v[1]) + f"  sum={tot/nsteps:.3f}")
    gp.close()


if __name__ == "__main__":
    which = sys.argv[1:] or ["fft", "rhs", "steps", "time"]
    def run(f, *a, **k):
        try:
            f(*a, **k)
        except Exception:
            print(f"!! {f.__name__}{a}{k} failed:"); traceback.print_exc()
    if "fft" in which:
        for dims in [(16, 16, 16), (32, 32, 32), (64, 32, 16), (16, 64, 128), (128, 128, 128), (256, 256, 256)]:
            run(fft_checks, *dims)
        run(fft_checks, 32, 32, 32, np.float64)
        run(fft_checks, 64, 128, 32, np.float64)
        run(fft_checks, 512, 512, 512)
    if "rhs" in which:
        for kind in ("hd", "mhd", "emhd"):
            run(rhs_check, kind, 32, turb=False)
            run(rhs_check, kind, 32, turb=True)
            run(rhs_check, kind, 32, T=np.float64, turb=True)
        run(rhs_check, "mhd", 0, turb=True, nxyz=(64, 32, 16))
        run(rhs_check, "mhd", 128, turb=True)
    if "steps" in which:
        for kind in ("hd", "mhd", "emhd"):
            run(step_check, kind, 32, stepper="RK4")
            run(step_check, kind, 32, stepper="LSRK54")
            run(step_check, kind, 32, T=np.float64, stepper="RK4")
        run(step_check, "mhd", 64, stepper="RK4", nsteps=6)
    if "time" in which:
        for n in (32, 128, 256, 512):
            run(timing, "mhd", n)
        run(timing, "hd", 256)
        run(timing, "mhd", 512, "LSRK54")
        run(timing, "emhd", 256)
        run(timing, "mhd", 1024, "RK4", 2)
