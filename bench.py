#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native MHDFlows hot path.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload mhd1024] [--scaling strong|weak]

A "step" is one RK4 (or LSRK54) step of the 3D pseudospectral problem on synthetic (analytic Taylor-Green) initial data.
Default workload: BASELINE.json's target configuration -- MHD Taylor-Green 1024^3 Float32 RK4 (configs[3]) -- on one GPU at
N = 1 and the SAME grid slab-decomposed over N GPUs (strong scaling) at N > 1.  The other configs (hd32, mhd256,
mhd512_lsrk, emhd512) are selected with --workload.  Prints ONE JSON line (rank 0).  DESIGN.md section 6 defines every key.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (kind, n, stepper, nu, eta, dt)
    "hd32": ("hd", 32, "RK4", 1e-2, 0.0, 1e-2),            # configs[0] (CPU-runnable reference case)
    "mhd256": ("mhd", 256, "RK4", 1e-3, 1e-3, 1e-3),        # configs[1]
    "mhd512_lsrk": ("mhd", 512, "LSRK54", 5e-4, 5e-4, 5e-4),  # configs[2]
    "mhd1024": ("mhd", 1024, "RK4", 2e-4, 2e-4, 2.5e-4),    # configs[3]  <- north_star target, the default
    "emhd512": ("emhd", 512, "RK4", 0.0, 0.0, 1e-5),        # configs[4]
    "mhd128": ("mhd", 128, "RK4", 2e-3, 2e-3, 2e-3),
    "mhd512": ("mhd", 512, "RK4", 5e-4, 5e-4, 5e-4),
}
DEFAULT_WORKLOAD = "mhd1024"
# SURVEY 8(d) contract figure: algorithmic bytes per step in units of S = 8 (N/2+1) N^2 bytes (NOT reduced for pruning)
ALG_S_PER_STEP = {("mhd", "RK4"): 384, ("mhd", "LSRK54"): 450, ("hd", "RK4"): 216, ("hd", "LSRK54"): 5 * 48,
                  ("emhd", "RK4"): 424, ("emhd", "LSRK54"): 5 * 100}
# the reference's only published number for this metric -- MHD Taylor-Green 256^3 Float32 RK4,
# 0.271 s per iteration on an RTX 3080 (README.md:78) = 6.19e7 grid-points*steps/s.  Nothing is published for other configs.
PUBLISHED_PTS_STEPS_PER_S = {"mhd256": 256 ** 3 / 0.271}
# fields through the fused x pass: (n_in, n_out); SURVEY 8(d): (n_in + n_out) S per launch
XPASS_FIELDS = {"mhd": (6, 9), "hd": (3, 6), "emhd": (18, 3)}
XPASS_S_PER_LAUNCH = {"mhd": 15, "hd": 9, "emhd": 19}
REF_SAMPLE_N = 128          # grid of the CPU-reference sample (full time steps of the same physics; see run_reference)


def grid_for(n, world):
    """Weak scaling: n^3 points per GPU.  The grid grows along z, then y, then x as ranks double
    (N=2: n x n x 2n, N=4: n x 2n x 2n, N=8: 2n x 2n x 2n); dx stays 2 pi / n so the physics per point is unchanged."""
    nx = ny = nz = n
    w = world
    for ax in ("z", "y", "x", "z", "y", "x"):
        if w <= 1:
            break
        if ax == "z":
            nz *= 2
        elif ax == "y":
            ny *= 2
        else:
            nx *= 2
        w //= 2
    return nx, ny, nz


def _tg_factors(n, dims, zrange):
    """1D factors of the analytic Taylor-Green u and b on x_i = -L/2 + i dx, dx = 2 pi / n, Float64.
    Every field is plane(y, x) * line(z); returns [(plane or None, line)] for ux, uy, uz, bx, by, bz."""
    nx, ny, nz = dims
    z0, z1 = zrange
    dx = 2 * math.pi / n
    X = (np.float32(-math.pi * nx / n) + np.float32(dx) * np.arange(nx)).astype(np.float64).reshape(1, -1)
    Y = (np.float32(-math.pi * ny / n) + np.float32(dx) * np.arange(ny)).astype(np.float64).reshape(-1, 1)
    Z = (np.float32(-math.pi * nz / n) + np.float32(dx) * np.arange(z0, z1)).astype(np.float64)
    return [(np.sin(X) * np.cos(Y), np.cos(Z)), (-np.cos(X) * np.sin(Y), np.cos(Z)), (None, None),
            (np.cos(X) * np.sin(Y), np.sin(Z)), (np.sin(X) * np.cos(Y), np.sin(Z)), (-2 * np.sin(X) * np.sin(Y), np.cos(Z))]


def tg_fields(n, T=np.float32, pinned=False, dims=None, zrange=None):
    """Analytic Taylor-Green u and b, shape (nz_local, ny, nx): [(array, owner)] for ux, uy, uz, bx, by, bz.
    Built plane by plane ((plane * line[z]) in Float64, rounded to T) on a thread pool: 1024^3 costs seconds and no
    Float64 3D temporaries (the six N^3 Float64 temporaries of round 1 cost 92 GPU-minutes of host time at 1024^3)."""
    nx, ny, nz = dims or (n, n, n)
    z0, z1 = zrange or (0, nz)
    shape = (z1 - z0, ny, nx)
    out = []
    zb = max(1, (1 << 22) // (nx * ny))          # z planes per task: ~32 MB Float64 temporaries
    pool = ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 1))
    for plane, line in _tg_factors(n, (nx, ny, nz), (z0, z1)):
        if pinned:
            import torch
            try:
                t = torch.empty(shape, dtype=torch.float32, pin_memory=True)
            except RuntimeError:      # a host that cannot pin this much memory: pageable buffers (slower e2e copies, same result)
                t = torch.empty(shape, dtype=torch.float32)
            a = t.numpy()
        else:
            t, a = None, np.empty(shape, dtype=T)
        if plane is None:
            a[...] = 0
        else:
            def fill(k0, a=a, plane=plane, line=line):
                k1 = min(k0 + zb, shape[0])
                a[k0:k1] = plane[None, :, :] * line[k0:k1, None, None]
            list(pool.map(fill, range(0, shape[0], zb)))
        out.append((a, t))
    pool.shutdown()
    return out


def tg_field_device(n, dims, idx, device):
    """Field `idx` of tg_fields on the whole grid as a CUDA tensor, bit-identical to the host version: the Float64 1D factors
    come from the host, the products (exactly rounded on both sides) are formed on the device.  No host 3D arrays."""
    import torch
    nx, ny, nz = dims
    plane, line = _tg_factors(n, dims, (0, nz))[idx]
    out = torch.zeros((nz, ny, nx), dtype=torch.float32, device=device)
    if plane is None:
        return out
    pl = torch.from_numpy(plane).to(device)
    ln = torch.from_numpy(line).to(device)
    zb = max(1, (1 << 25) // (nx * ny))
    for k0 in range(0, nz, zb):
        k1 = min(k0 + zb, nz)
        out[k0:k1] = (pl[None, :, :] * ln[k0:k1, None, None]).to(torch.float32)
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device=0):
        self.device, self.rows, self._stop, self._th = device, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                r = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.device)],
                                   capture_output=True, text=True, timeout=5)
                if r.returncode == 0 and r.stdout.strip():
                    self.rows.append([c.strip() for c in r.stdout.strip().split(",")])
            except Exception:
                pass
            self._stop.wait(0.1)

    def start(self):
        self._th = threading.Thread(target=self._run, daemon=True)
        self._th.start()

    def stop(self):
        self._stop.set()
        if self._th:
            self._th.join(timeout=6)
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                smax.append(float(r[1]))
                for nm, v in zip(names, r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def make_problem(kind, n, stepper, nu, eta, dt, device=0, dims=None, rank=0, world=1, nccl_id=None):
    import mhdflows_jl_b200 as M
    nx, ny, nz = dims or (n, n, n)
    L0 = 2 * math.pi
    kw = dict(nx=nx, ny=ny, nz=nz, Lx=L0 * nx / n, Ly=L0 * ny / n, Lz=L0 * nz / n, nu=nu, eta=eta, dt=dt, stepper=stepper)
    if kind == "mhd":
        kw.update(B_field=True)
    elif kind == "emhd":
        kw.update(B_field=True, EMHD=True)
    if world > 1:
        kw.update(rank=rank, nranks=world, nccl_id=nccl_id)
    p = M.Problem(M.GPU(device), **kw)
    if world > 1:
        from mhdflows_jl_b200.dist import enable_peer_exchange
        enable_peer_exchange(p)
    return M, p


def nid2(world):
    if world <= 1:
        return None
    from mhdflows_jl_b200.dist import nccl_id_via_torch
    return nccl_id_via_torch()


def set_ic(M, p, kind, fields):
    if kind == "emhd":
        M.SetUpProblemIC(p, bx=fields[3][0], by=fields[4][0], bz=fields[5][0])
    elif kind == "mhd":
        M.SetUpProblemIC(p, ux=fields[0][0], uy=fields[1][0], uz=fields[2][0], bx=fields[3][0], by=fields[4][0], bz=fields[5][0])
    else:
        M.SetUpProblemIC(p, ux=fields[0][0], uy=fields[1][0], uz=fields[2][0])


def field_names(kind):
    return ["bx", "by", "bz"] if kind == "emhd" else (["ux", "uy", "uz", "bx", "by", "bz"] if kind == "mhd" else ["ux", "uy", "uz"])


def oracle_problem(kind, n, stepper, nu, eta, dt):
    from oracle import mhdflows_oracle as O
    kw = dict(nx=n, T=np.float32, dt=dt, stepper=stepper)
    if kind == "mhd":
        op = O.Problem(nu=nu, eta=eta, B_field=True, **kw)
    elif kind == "emhd":
        op = O.Problem(B_field=True, EMHD=True, **kw)
    else:
        op = O.Problem(nu=nu, **kw)
    ic = O.taylor_green_ic(op.grid)
    if kind == "emhd":
        O.SetUpProblemIC(op, bx=ic[3], by=ic[4], bz=ic[5])
    elif kind == "mhd":
        O.SetUpProblemIC(op, *ic[:3], bx=ic[3], by=ic[4], bz=ic[5])
    else:
        O.SetUpProblemIC(op, *ic[:3])
    return O, op


def cpu_oracle_steps(kind, n, stepper, nu, eta, dt, nsteps, warm):
    """Time FULL time steps of the CPU oracle (literal restatement of the reference op sequence: 36 / 24 / 51 3D FFTs per RHS
    evaluation, 4 or 5 evaluations per step; scipy pocketfft on all host threads).  Returns seconds per timed step."""
    O, op = oracle_problem(kind, n, stepper, nu, eta, dt)
    times = []
    for i in range(warm + nsteps):
        t0 = time.perf_counter()
        O.stepforward(op)
        t1 = time.perf_counter()
        if i >= warm:
            times.append(t1 - t0)
    return times


def ref_sample(wl):
    """The CPU-reference sample of a workload: the same physics / stepper / parameters on a grid the CPU finishes in seconds
    per step (the metric, grid-points*steps/s, is normalised by the grid size)."""
    kind, n, stepper, nu, eta, dt = WORKLOADS[wl]
    ns = min(n, REF_SAMPLE_N)
    text = (f"full {stepper} steps of the {kind.upper()} oracle port ({ {'mhd': 36, 'hd': 24, 'emhd': 51}[kind]} 3D FFTs per RHS evaluation) "
            f"on a {ns}^3 Taylor-Green grid" + ("" if ns == n else f" (bounded sample of {wl}: {n}^3 does not fit the time limit on a CPU; "
                                                 "value = sample grid points * steps / s)"))
    return (kind, ns, stepper, nu, eta, dt), text


def config_for(wl, dims, world, strong):
    """The `config` object -- identical in both arms (--impl ours / reference) for the same command line."""
    kind, n, stepper, nu, eta, dt = WORKLOADS[wl]
    return {"workload": wl, "kind": kind, "n": n, "grid": list(dims), "stepper": stepper, "nu": nu, "eta": eta, "dt": dt,
            "ic": "analytic Taylor-Green u and b",
            "l2": "per-step working set (FFT work buffers, >= 1.4 GB at 256^3) is far larger than the 126 MB L2; no flush needed",
            "multi_gpu": "single" if world == 1 else ("slab decomposition (z slabs / ky slabs) of the fixed grid" if strong
                                                       else f"slab decomposition, {n}^3 points per GPU")}


def cufft_reference_point(kind, dims, reps=3):
    """cuFFT timed alongside (north_star): the 3D transforms of one RHS evaluation through torch.fft (= cuFFT plans), without
    any of the products / spectral work -- (a) the fused formulation's count (MHD 6 c2r + 9 r2c), (b) the reference's own
    count (MHD 36).  A library reference point, never the product path."""
    try:
        import torch
        nx, ny, nz = dims
        n_c2r, n_r2c = XPASS_FIELDS[kind]
        if kind == "emhd":
            n_c2r = 24          # row transforms of the gradient form (18 fields through the y / z passes)
        ref_total = {"mhd": 36, "hd": 24, "emhd": 51}[kind]
        x = torch.randn((nz, ny, nx), device="cuda", dtype=torch.float32)
        xh = torch.fft.rfftn(x)
        for _ in range(2):
            torch.fft.irfftn(xh, s=x.shape)
            torch.fft.rfftn(x)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record()
        for _ in range(reps):
            torch.fft.irfftn(xh, s=x.shape)
        ev[1].record()
        for _ in range(reps):
            torch.fft.rfftn(x)
        ev[2].record()
        torch.cuda.synchronize()
        c2r_ms, r2c_ms = ev[0].elapsed_time(ev[1]) / reps, ev[1].elapsed_time(ev[2]) / reps
        fused = n_c2r * c2r_ms + n_r2c * r2c_ms
        n_c2r_ref = {"mhd": 6, "hd": 3, "emhd": 7}[kind]
        literal = n_c2r_ref * c2r_ms + (ref_total - n_c2r_ref) * r2c_ms
        del x, xh
        torch.cuda.empty_cache()
        return {"library": "cuFFT via torch.fft (out-of-place, full (N/2+1)N^2 spectra, no pruning)", "c2r_ms": c2r_ms, "r2c_ms": r2c_ms,
                "ffts_per_rhs_fused_form": n_c2r + n_r2c, "fft_only_ms_per_rhs_fused_form": fused,
                "ffts_per_rhs_reference_form": ref_total, "fft_only_ms_per_rhs_reference_form": literal,
                "what": "transforms only; our ms_per_step / stages also contains the products, the spectral assembly and the RK update"}
    except Exception as e:      # a reference point must never take the bench line down
        return {"unavailable": f"{type(e).__name__}: {e}"[:200]}


def gpu_baseline_point(kind, n, nu, eta, dt, stepper="RK4", steps=3, warm=1):
    """`gpu_baseline`: the reference's own GPU formulation (CUDA.jl path: 36-FFT literal op sequence, cuFFT + one unfused kernel
    per broadcast, FourierFlows RK4 with sol1 + 4 RHS arrays) restated in torch eager mode (baseline/torch_v0.py), timed on
    this GPU beside the product.  Its layout (36 S + 7 R of state and scratch) does not fit 1024^3 in 180 GB -- the reference
    README tops out at 700^3 on 80 GB -- so grids above 512^3 are timed at 512^3 (the metric is per grid point)."""
    try:
        import torch
        from baseline.torch_v0 import TorchV0
        if kind == "emhd":
            return {"unavailable": "the torch stand-in covers the HD / MHD RHS only"}
        nb = min(n, 512)
        b = TorchV0(nb, kind=kind, nu=nu, eta=eta, dt=dt, stepper=stepper)
        for i in range(b.Nl):       # IC field by field on the device
            f = tg_field_device(nb, (nb, nb, nb), i, "cuda")
            b.vars[i].copy_(f)
            b.sol[i] = b.rfft(b.vars[i])
            del f
        for _ in range(warm):
            b.stepforward()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            b.stepforward()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        mem = torch.cuda.max_memory_allocated()
        del b
        torch.cuda.empty_cache()
        return {"value": nb ** 3 / (ms * 1e-3), "unit": "pts*steps/s", "ms_per_step": ms, "grid": [nb, nb, nb], "steps": steps, "warmup": warm,
                "kind": "port", "peak_bytes": int(mem),
                "what": "reference op sequence (MHDSolver.jl:330-351: 36 cuFFT transforms per RHS + one eager kernel per broadcast, "
                        f"FourierFlows {stepper}) in torch on this GPU; stand-in for the CUDA.jl path, which cannot be installed offline"}
    except Exception as e:
        return {"unavailable": f"{type(e).__name__}: {e}"[:300]}


def pruned_bytes(kind, stepper, info, dims, world, F):
    """Bytes the pruned layout must move per step and per x-pass launch, per GPU (every pass = read input + write output once;
    compact field cf = Kxp Kyl Kz, after z pass nz Kyl Kxp, x-pass layout nzl ny Kxp; 8 bytes per Float32 complex)."""
    nx, ny, nz = dims
    Kxp, Ky, Kz = info["Kxp"], info["Ky"], info["Kz"]
    Kyl = -(-Ky // world)
    nzl = nz // world
    cf, zf, xf = Kxp * Kyl * Kz * 8, nz * Kyl * Kxp * 8, nzl * ny * Kxp * 8
    nin, nout = XPASS_FIELDS[kind]
    xpass = (nin + nout) * xf + (6 * nx * ny * nzl * 4 if kind == "emhd" else 0)
    stage = nin * (cf + zf) + nin * (zf + xf) + xpass + nout * (xf + zf) + nout * (zf + cf)
    if stepper == "RK4":      # spectral update over the 4 stages: reads P x4, Sin x4, Y x2 (stages 2, 3), A x3; writes A x3, Sout x4
        spec = (4 * nout + 16 * F) * cf
        nst = 4
    else:
        spec = 5 * (nout + 4 * F) * cf
        nst = 5
    if kind == "emhd":
        stage += (3 + 18) * cf            # k_emhd_derive
    return nst * stage + spec, xpass


def run_reference(args, wl, dims, world, strong):
    """--impl reference: the reference's CPU path on the host cores.  Julia + FFTW cannot run here (not installed, no network;
    DESIGN.md section 2), so the oracle port (same 36-FFT op sequence, scipy pocketfft on all host threads) is timed on a
    bounded sample: every timed step is one FULL time step of the same physics on a REF_SAMPLE_N^3 grid."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    (kind, ns, stepper, nu, eta, dt), text = ref_sample(wl)
    cores = os.cpu_count() or 1
    times = cpu_oracle_steps(kind, ns, stepper, nu, eta, dt, nsteps=args.steps, warm=args.warmup)
    sec = float(sum(times)) / len(times)
    val = ns ** 3 / sec
    line = {"impl": "reference", "metric": "grid-points*steps/s", "value": val, "unit": "pts*steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_for(wl, dims, world, strong),
            "cpu_baseline": {"value": val, "unit": "pts*steps/s", "cores": cores, "kind": "port", "sample": text,
                             "note": "reference CPU path = oracle port (NumPy + scipy.fft pocketfft); Julia / FFTW are not installable offline"},
            "e2e": {"value": val, "unit": "pts*steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def parity_protocol(M, p, kind, dt, fields=None, device_fields=None, nsteps=2):
    """Fixed protocol of the multi-GPU parity check: Taylor-Green IC, `nsteps` steps of the workload's dt, then the fresh
    energies, the three helicities and the shell spectra of every state field."""
    names = field_names(kind)
    if fields is not None:
        set_ic(M, p, kind, fields)
    else:
        for nm in names:
            f = device_fields(nm)
            p.set_real(nm, f)
            del f
    p.clock.dt = dt
    p.clock.t = 0.0
    M.stepforward(p, nsteps)
    ke, me = p.energy(M.FRESH)
    hk, hm, hc = p.helicity()
    spec = np.stack([M.spectralline(p, i, nbins=64)[0].astype(np.float64) for i in range(len(names))])
    return {"KE": ke, "ME": me, "Hk": hk, "Hm": hm, "Hc": hc, "spectrum_checksum": float(spec.sum()), "spectrum": spec}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--scaling", default=None, choices=["weak", "strong"],
                    help="N > 1: strong = the workload's grid split over N GPUs (default), weak = the workload's n^3 points per GPU")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    wl = args.workload or DEFAULT_WORKLOAD
    if wl not in WORKLOADS:
        raise SystemExit(f"unknown workload {wl}; choose from {sorted(WORKLOADS)}")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.scaling is None:
        args.scaling = "strong" if args.gpus > 1 else "weak"      # N = 1: one GPU's work is fixed either way
    strong = args.scaling == "strong"
    kind, n, stepper, nu, eta, dt = WORKLOADS[wl]
    dims = (n, n, n) if (args.gpus == 1 or strong) else grid_for(n, args.gpus)
    if args.impl == "reference":
        run_reference(args, wl, dims, args.gpus, strong)
        return
    if world != args.gpus and args.gpus != 1:
        raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N > 1")
    import torch
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    K, W = args.steps, args.warmup
    nid = None
    if world > 1:
        from mhdflows_jl_b200.dist import nccl_id_via_torch
        nid = nccl_id_via_torch()
    t_setup0 = time.perf_counter()
    M, p = make_problem(kind, n, stepper, nu, eta, dt, device=local, dims=dims, rank=rank, world=world, nccl_id=nid)
    nzl = dims[2] // world
    fields = tg_fields(n, pinned=True, dims=dims, zrange=(rank * nzl, (rank + 1) * nzl))
    set_ic(M, p, kind, fields)
    setup_s = time.perf_counter() - t_setup0
    S = 8 * (dims[0] // 2 + 1) * dims[1] * dims[2]      # one reference-layout spectral field of the whole grid
    npts = dims[0] * dims[1] * dims[2]
    info = p.info()
    F = info["nfields"]

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timed region: `value` ------------------------------------------------
    p.step_timed(W)
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    l0 = p.launch_count()
    ms = p.step_timed(K)
    l1 = p.launch_count()
    barrier()
    if dist is not None:
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / K
    value = npts * K / (ms * 1e-3)           # whole grid (all ranks) per second

    # ---- second pass with per-launch CUDA events: roofline of the dominant kernel ----------
    p.profile(True)
    ms_prof = p.step_timed(K)
    prof = p.profile_get()
    p.profile(False)
    clocks = sampler.stop()
    peak, peak_src = measured_peak_hbm()
    x_ms, x_cnt = prof["x_fused"]
    x_avg_s = x_ms * 1e-3 / max(x_cnt, 1)
    x_bytes = XPASS_S_PER_LAUNCH[kind] * S // world   # contract figure, this rank's share of the rows
    achieved = x_bytes / x_avg_s / 1e9
    alg_step = ALG_S_PER_STEP[(kind, stepper)] * S
    step_ach = alg_step / world / (ms_per_step * 1e-3) / 1e9   # per GPU
    pr_step, pr_x = pruned_bytes(kind, stepper, info, dims, world, F)
    stages = 4 if stepper == "RK4" else 5
    x_launches_per_stage = max(1.0, x_cnt / (K * stages))      # z-chunk pipelined slab runs launch the x pass once per chunk
    pr_x_launch = pr_x / x_launches_per_stage
    x_bytes_launch = x_bytes / x_launches_per_stage
    achieved = x_bytes_launch / x_avg_s / 1e9
    comp_total = sum(v[0] for k_, v in prof.items() if v[1] and k_ != "exchange")
    shares = {k_: v[0] / max(comp_total, 1e-12) for k_, v in prof.items() if v[1] and k_ != "exchange"}
    class_ms = {k_: v[0] / K for k_, v in prof.items() if v[1]}
    nvlink = None
    if world > 1:
        # NVLink roofline of the global transposes: the bytes the pruned exchange really pushes per GPU per step over
        # the time the exchange is in flight on the communication stream (CUDA events; includes the cross-rank flags / barriers);
        # peak = 900 GB/s per direction per GPU.  `exposed_ms_per_step` = step time minus the sum of this rank's compute
        # kernels (both from the profiled pass): the part of the exchange that is NOT hidden behind the FFT passes.
        try:
            nf_x = sum(XPASS_FIELDS[kind])
            contract = stages * nf_x * (S / world) * (world - 1) / world
            real = stages * nf_x * (dims[2] // world) * p.layout.Kyl * info["Kxp"] * 8 * (world - 1)
            ex_ms, ex_cnt = prof["exchange"]
            ex_s_per_step = ex_ms * 1e-3 / K
            nvlink = {"peak": 900.0, "unit": "GB/s", "contract_bytes_per_step_per_gpu": contract, "pushed_bytes_per_step_per_gpu": real,
                      "exchange_in_flight_ms_per_step": ex_s_per_step * 1e3, "exchanges_per_step": ex_cnt / K,
                      "achieved_pushed": real / ex_s_per_step / 1e9, "frac_pushed": real / ex_s_per_step / 1e9 / 900.0,
                      "exposed_ms_per_step": max(0.0, ms_prof / K - comp_total / K),
                      "compute_ms_per_step": comp_total / K}
        except Exception as e:
            nvlink = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
    traffic = None
    tp = os.path.join(ROOT, "profiles", "xfused_traffic.json")
    if os.path.exists(tp):      # dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture, per workload / GPU count
        try:
            with open(tp) as f:
                traffic = json.load(f).get(wl if world == 1 else f"{wl}@{world}")
        except Exception:
            traffic = None

    # ---- end-to-end through the public API with host buffers: `e2e` ----------------------
    # ONE timed region per run = SetUpProblemIC! from pinned host arrays (H2D of every field) + K iterations of the
    # reference's TimeIntegrator! loop body (getCFL! -> stepforward! -> ProbDiagnostic, scalars D2H every step) + a
    # savefile-style download of every real field into pinned host arrays (D2H).  Bytes are per run / K.
    nf = len(field_names(kind))
    p.close()
    M2, q = make_problem(kind, n, stepper, nu, eta, dt, device=local, dims=dims, rank=rank, world=world, nccl_id=nid2(world))
    set_ic(M2, q, kind, fields)
    M2.stepforward(q, 1)
    names = field_names(kind)
    src = [fields[{"ux": 0, "uy": 1, "uz": 2, "bx": 3, "by": 4, "bz": 5}[nm]][0] for nm in names]
    barrier()
    t0 = time.perf_counter()
    set_ic(M2, q, kind, fields)
    g = q.grid
    dl = min(g.Lx / g.nx, g.Ly / g.ny, g.Lz / g.nz)
    vi = max(nu, eta)
    t_diff = 0.25 * dl * dl / vi if vi > 0 else math.inf
    for _ in range(K):
        M2.getCFL(q, t_diff, Coef=0.25)
        M2.stepforward(q)
        M2.ProbDiagnostic(q)
    for nm, buf in zip(names, src):          # the pinned upload buffers receive the downloaded fields (inputs are consumed)
        q.get_real(nm, M2.STALE, out=buf)
    barrier()
    t1 = time.perf_counter()
    e2e_s = t1 - t0
    if dist is not None:
        t = torch.tensor([e2e_s], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_val = npts * K / e2e_s
    h2d = nf * npts * 4 / K          # all ranks together
    d2h = nf * npts * 4 / K + 88 * world

    # ---- multi-GPU parity: the N-rank run against a single-GPU run of the same grid --------------------
    parity = None
    if world > 1 and not args.no_parity:
        try:
            fields = tg_fields(n, pinned=True, dims=dims, zrange=(rank * nzl, (rank + 1) * nzl))   # the downloads overwrote them
            mine = parity_protocol(M2, q, kind, dt, fields=fields)
            del fields
            q.close()
            q = None
            barrier()
            if rank == 0:
                M3, s = make_problem(kind, n, stepper, nu, eta, dt, device=local, dims=dims)
                idx = {"ux": 0, "uy": 1, "uz": 2, "bx": 3, "by": 4, "bz": 5}
                ref = parity_protocol(M3, s, kind, dt, device_fields=lambda nm: tg_field_device(n, dims, idx[nm], f"cuda:{local}"))
                s.close()
                rel = {k_: abs(mine[k_] - ref[k_]) / max(abs(ref[k_]), 1e-300) for k_ in ("KE", "ME", "Hk", "Hm", "Hc", "spectrum_checksum")}
                # helicities of the Taylor-Green fields are zero by symmetry: compare those on the scale of the energies
                scale = max(abs(ref["KE"]), abs(ref["ME"]))
                for k_ in ("Hk", "Hm", "Hc"):
                    rel[k_] = abs(mine[k_] - ref[k_]) / max(abs(ref[k_]), scale)
                sp = float(np.max(np.abs(mine["spectrum"] - ref["spectrum"]) / np.maximum(np.abs(ref["spectrum"]), 1e-30 + 1e-12 * ref["spectrum"].max())))
                parity = {"protocol": "Taylor-Green IC, 2 steps of the workload dt, then fresh KE / ME, helicities, 64-bin shell spectra of every field",
                          "reference": "single-GPU run of the same grid in the same job (rank 0)",
                          "n_rank": {k_: mine[k_] for k_ in rel}, "single_gpu": {k_: ref[k_] for k_ in rel},
                          "rel_diff": rel, "spectrum_bins_max_rel_diff": sp, "max_rel_diff": max(max(rel.values()), sp),
                          "ok": bool(max(max(rel.values()), sp) <= 1e-10)}
        except Exception as e:
            parity = {"unavailable": f"{type(e).__name__}: {e}"[:300]}
    if q is not None:
        q.close()
    del src
    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return
    cfg = config_for(wl, dims, world, strong)
    cfg["baseline"] = ("MHDFlows.jl README.md:78: 0.271 s/iteration, MHD TG 256^3 Float32 RK4 on an RTX 3080 (CUDA.jl)"
                       if wl in PUBLISHED_PTS_STEPS_PER_S else None)
    line = {
        "metric": "grid-points*steps/s", "value": value, "unit": "pts*steps/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": (value / PUBLISHED_PTS_STEPS_PER_S[wl]) if (world == 1 and wl in PUBLISHED_PTS_STEPS_PER_S) else None,
        "dtype": "f32", "data": "synthetic",
        "config": cfg,
        "roofline": {"bound": "hbm", "kernel": "k_xfused (c2r -> products -> r2c)", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": x_bytes_launch,
                     "algorithmic_bytes_note": "SURVEY 8(d) contract figure (n_in + n_out) S over unpruned (N/2+1) N^2 rows",
                     "pruned": {"bytes_per_launch": pr_x_launch, "achieved": pr_x_launch / x_avg_s / 1e9, "frac": pr_x_launch / x_avg_s / 1e9 / peak,
                                "note": "bytes the dealiased-band layout has to move (x rows hold Kxp of N/2+1 columns): the honest HBM fraction of this kernel"},
                     "avg_launch_ms": x_avg_s * 1e3, "launches_timed": int(x_cnt),
                     "timed_in": "second pass of K steps with a CUDA event pair around every launch on the library stream",
                     "kernel_share_of_compute": shares, "class_ms_per_step": class_ms,
                     "step": {"contract_bytes": alg_step, "contract_ratio": step_ach / peak,
                              "contract_note": "384 S-type figure of SURVEY 8(d), not reduced for pruning: a ratio (may exceed 1), not a roofline fraction",
                              "pruned_bytes_per_gpu": pr_step, "pruned_achieved": pr_step / (ms_per_step * 1e-3) / 1e9,
                              "frac_pruned": pr_step / (ms_per_step * 1e-3) / 1e9 / peak,
                              "ms_per_step_profiled": ms_prof / K}},
        "e2e": {"value": e2e_val, "unit": "pts*steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_s / K * 1e3,
                "what": "per run / K: ONE SetUpProblemIC! from pinned host + K x (getCFL!, stepforward!, ProbDiagnostic) + ONE download of all real fields"},
        "gpu_launches": int(l1 - l0),
        "clocks": clocks,
        "setup_s": setup_s,
    }
    if nvlink is not None:
        line["nvlink"] = nvlink
    if parity is not None:
        line["parity"] = parity
    if world == 1:
        cf = cufft_reference_point(kind, dims)
        if cf and "fft_only_ms_per_rhs_fused_form" in cf:
            cf["ours_ms_per_rhs_everything_included"] = ms_per_step / stages
        line["cufft_ref"] = cf
        if not args.no_gpu_baseline:
            gb = gpu_baseline_point(kind, n, nu, eta, dt, stepper=stepper)
            if "value" in gb:
                gb["ours_over_gpu_baseline"] = value / gb["value"]
            line["gpu_baseline"] = gb
    if not args.no_cpu_baseline and world == 1:
        (k2, ns, st2, nu2, eta2, dt2), text = ref_sample(wl)
        times = cpu_oracle_steps(k2, ns, st2, nu2, eta2, dt2, nsteps=4, warm=1)
        sec = float(sum(times)) / len(times)
        line["cpu_baseline"] = {"value": ns ** 3 / sec, "unit": "pts*steps/s", "cores": os.cpu_count() or 1, "kind": "port",
                                "sample": text + "; 1 warm-up + 4 timed steps", "ms_per_step": sec * 1e3}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
