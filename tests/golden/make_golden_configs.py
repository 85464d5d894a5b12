"""Golden digests of the CPU oracle at the BASELINE.json configuration sizes (configs 2, 3, 5) -- too slow to recompute inside
the GPU suite (minutes to an hour of pocketfft), too large to commit as full states, so what is frozen is a digest:

    python tests/golden/make_golden_configs.py [cfg3 cfg5 cfg2]       (about an hour on 8 cores for all three)

  cfg3  MHD decaying turbulence 512^3 Float32 LSRK54, random-phase IC (DivFreeSpectraMap, k0 = -5/6, P = 1, Philox phases with
        seeds 1234 / 5678 -- the stream the device generates, regenerated here in NumPy), nu = eta = 5e-4, dt = 5e-4:
        N = calcN!(sol0) and sol after 2 steps
  cfg5  EMHD 512^3 Float32 RK4, Taylor-Green b: N = calcN!(sol0) and sol after 1 step of dt = 1e-5
  cfg2  MHD Taylor-Green 256^3 Float32 RK4, nu = eta = 1e-3, dt = 1e-3: KE / ME of the stale vars after every one of 100
        steps, helicities every 10th step, sol after 100 steps
Each digest holds the values at 20000 randomly chosen retained modes per field group, the L2 norms over all retained modes
and the shell spectra (spectralline) of every field.  Like every golden file of this repo they pin the ORACLE (parity
unpinned: the reference has no fixtures and Julia cannot run here).  tests/test_gpu_configs.py compares the CUDA path.
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import forcing_oracle as FO  # noqa: E402
from oracle import mhdflows_oracle as O  # noqa: E402

DFSM_CALL = 0x7FFFFFFF44465350      # counter tag of the device random-phase stream (mhdflows_jl_b200.DFSM_CALL)
NSAMPLE = 20000
NPEAK = 256


def sample_index(g, nfields, seed=0):
    """Flat indices of NSAMPLE retained modes per field (the same for every digest of a grid)."""
    flat = np.flatnonzero(g.retained_mask().ravel())
    rng = np.random.default_rng(seed)
    return np.stack([rng.choice(flat, NSAMPLE, replace=False) for _ in range(nfields)])


def digest(g, arr, idx, prefix, out):
    """arr: (F, nz, ny, nkr) spectral fields (dealiased copy is taken)."""
    a = g.dealias(arr.copy())
    out[prefix + "_samples"] = np.stack([a[f].ravel()[idx[f]] for f in range(a.shape[0])])
    out[prefix + "_norms"] = np.array([float(np.linalg.norm(a[f].astype(np.complex128).ravel())) for f in range(a.shape[0])])
    out[prefix + "_spectra"] = np.stack([O.spectralline(g.irfft(a[f].copy()), g)[0].astype(np.float64) for f in range(a.shape[0])])
    # the NPEAK largest modes of every field: on sparse (Taylor-Green) spectra they ARE the signal, the rest is rounding noise
    pk_i, pk_v = [], []
    for f in range(a.shape[0]):
        flat = a[f].ravel()
        i = np.argpartition(np.abs(flat), -NPEAK)[-NPEAK:]
        i = i[np.argsort(-np.abs(flat[i]))]
        pk_i.append(i)
        pk_v.append(flat[i])
    out[prefix + "_peak_index"] = np.stack(pk_i)
    out[prefix + "_peak_values"] = np.stack(pk_v)


def cfg3():
    kind, n, stepper, nu, eta, dt = "mhd", 512, "LSRK54", 5e-4, 5e-4, 5e-4
    op = O.Problem(nx=n, T=np.float32, nu=nu, eta=eta, dt=dt, stepper=stepper, B_field=True)
    g = op.grid
    u = O.DivFreeSpectraMap(g, FO.PhiloxField(1234, g).uniforms(DFSM_CALL)[0], k_peak=0.0, P=1, k0=-5 / 6)
    b = O.DivFreeSpectraMap(g, FO.PhiloxField(5678, g).uniforms(DFSM_CALL)[0], k_peak=0.0, P=1, k0=-5 / 6)
    O.SetUpProblemIC(op, *u, bx=b[0], by=b[1], bz=b[2])
    del u, b
    out = dict(n=n, nu=nu, eta=eta, dt=dt, seeds=np.array([1234, 5678]), k0=-5 / 6)
    idx = sample_index(g, 6)
    out["index"] = idx
    digest(g, op.sol, idx, "sol0", out)
    N = np.zeros_like(op.sol)
    op.calcN(N, op.sol.copy(), 0.0, op.clock, op.vars, op.params, g)
    digest(g, N, idx, "N", out)
    del N
    for _ in range(2):
        O.stepforward(op)
    digest(g, op.sol, idx, "sol2", out)
    out["energy_stale"] = np.array(O.ProbDiagnostic(op, rounded=False))
    np.savez_compressed(os.path.join(HERE, "cfg3_mhd512_lsrk54.npz"), **out)


def cfg5():
    n, dt = 512, 1e-5
    op = O.Problem(nx=n, T=np.float32, dt=dt, stepper="RK4", B_field=True, EMHD=True)
    g = op.grid
    ic = O.taylor_green_ic(g)
    O.SetUpProblemIC(op, bx=ic[3], by=ic[4], bz=ic[5])
    del ic
    out = dict(n=n, dt=dt)
    idx = sample_index(g, 3)
    out["index"] = idx
    N = np.zeros_like(op.sol)
    vars0 = {k: getattr(op.vars, k).copy() for k in ("bx", "by", "bz")}
    op.calcN(N, op.sol.copy(), 0.0, op.clock, op.vars, op.params, g)
    digest(g, N, idx, "N", out)
    del N
    for k, v in vars0.items():        # calcN! refreshed the stale b: start the step from the state SetUpProblemIC! left
        getattr(op.vars, k)[...] = v
    O.stepforward(op)
    digest(g, op.sol, idx, "sol1", out)
    out["energy_stale"] = np.array([O.ProbDiagnostic(op, rounded=False)])
    np.savez_compressed(os.path.join(HERE, "cfg5_emhd512_rk4.npz"), **out)


def cfg2():
    n, nu, eta, dt, nsteps = 256, 1e-3, 1e-3, 1e-3, 100
    op = O.Problem(nx=n, T=np.float32, nu=nu, eta=eta, dt=dt, stepper="RK4", B_field=True)
    g = op.grid
    ic = O.taylor_green_ic(g)
    O.SetUpProblemIC(op, *ic[:3], bx=ic[3], by=ic[4], bz=ic[5])
    del ic
    dV = g.dx * g.dy * g.dz
    ke, me, hel = [], [], []
    for s in range(nsteps):
        O.stepforward(op)
        KE, ME = O.ProbDiagnostic(op, rounded=False)
        ke.append(KE)
        me.append(ME)
        if s % 10 == 9:
            u = [g.irfft(g.dealias(op.sol[i].copy())) for i in range(3)]
            b = [g.irfft(g.dealias(op.sol[3 + i].copy())) for i in range(3)]
            Hk = float(np.sum(O.h_k(*u, g).astype(np.float64)))
            Hm = float(np.sum(O.h_m(*b, g).astype(np.float64)))
            Hc = float(sum(np.sum(a.astype(np.float64) * c) for a, c in zip(u, b))) * dV
            hel.append((s + 1, Hk, Hm, Hc))
            print(f"cfg2 step {s + 1}: KE {KE:.6f} ME {ME:.6f} Hk {Hk:.3e} Hm {Hm:.3e} Hc {Hc:.3e}", flush=True)
    out = dict(n=n, nu=nu, eta=eta, dt=dt, KE=np.array(ke), ME=np.array(me), helicity=np.array(hel))
    idx = sample_index(g, 6)
    out["index"] = idx
    digest(g, op.sol, idx, "sol100", out)
    np.savez_compressed(os.path.join(HERE, "cfg2_mhd256_rk4_100steps.npz"), **out)


if __name__ == "__main__":
    which = sys.argv[1:] or ["cfg3", "cfg5", "cfg2"]
    for w in which:
        t0 = time.time()
        {"cfg3": cfg3, "cfg5": cfg5, "cfg2": cfg2}[w]()
        print(f"{w}: done in {time.time() - t0:.0f} s", flush=True)
