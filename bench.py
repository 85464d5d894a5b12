#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native MHDFlows hot path.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload mhd256]

A "step" is one RK4 (or LSRK54) step of the 3D pseudospectral problem on synthetic (analytic Taylor-Green)
initial data.  Default workload at N=1: BASELINE.json configs[1] -- MHD Taylor-Green 256^3 Float32 RK4.
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for the definitions of every key.
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

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (kind, n, stepper, nu, eta, dt)
    "hd32": ("hd", 32, "RK4", 1e-2, 0.0, 1e-2),            # configs[0] (CPU-runnable reference case)
    "mhd256": ("mhd", 256, "RK4", 1e-3, 1e-3, 1e-3),        # configs[1]  <- the metric's single-GPU config
    "mhd512_lsrk": ("mhd", 512, "LSRK54", 5e-4, 5e-4, 5e-4),  # configs[2]
    "mhd1024": ("mhd", 1024, "RK4", 2e-4, 2e-4, 2.5e-4),    # configs[3]
    "emhd512": ("emhd", 512, "RK4", 0.0, 0.0, 1e-5),        # configs[4]
    "mhd128": ("mhd", 128, "RK4", 2e-3, 2e-3, 2e-3),
    "mhd512": ("mhd", 512, "RK4", 5e-4, 5e-4, 5e-4),
}
# algorithmic bytes per step in units of S = 8 (N/2+1) N^2 bytes
ALG_S_PER_STEP = {("mhd", "RK4"): 384, ("mhd", "LSRK54"): 450, ("hd", "RK4"): 216, ("hd", "LSRK54"): 5 * 48,
                  ("emhd", "RK4"): 424, ("emhd", "LSRK54"): 5 * 100}
# the reference's only published number for this metric and config -- MHD Taylor-Green 256^3
# Float32 RK4, 0.271 s per iteration on an RTX 3080 (README.md:78) = 6.19e7 grid-points*steps/s.  Other hardware; quoted
# because it is the only published figure.  No published number exists for any other workload or for N > 1.
PUBLISHED_PTS_STEPS_PER_S = {"mhd256": 256 ** 3 / 0.271}
# x-pass share of the model: (n_in + n_out) S per launch (one launch per stage)
XPASS_S_PER_LAUNCH = {"mhd": 15, "hd": 9, "emhd": 19}


def grid_for(n, world):
    """Weak scaling: 256^3-type workload per GPU.  The grid grows along z, then y, then x as ranks double
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


def tg_fields(n, T=np.float32, pinned=False, dims=None, zrange=None):
    """Analytic Taylor-Green u and b on x_i = -L/2 + i dx, dx = 2 pi / n, shape (nz_local, ny, nx)."""
    nx, ny, nz = dims or (n, n, n)
    z0, z1 = zrange or (0, nz)
    dx = 2 * math.pi / n
    X = (np.float32(-math.pi * nx / n) + np.float32(dx) * np.arange(nx)).astype(np.float64).reshape(1, 1, -1)
    Y = (np.float32(-math.pi * ny / n) + np.float32(dx) * np.arange(ny)).astype(np.float64).reshape(1, -1, 1)
    Z = (np.float32(-math.pi * nz / n) + np.float32(dx) * np.arange(z0, z1)).astype(np.float64).reshape(-1, 1, 1)
    shape = (z1 - z0, ny, nx)
    fs = [np.sin(X) * np.cos(Y) * np.cos(Z), -np.cos(X) * np.sin(Y) * np.cos(Z), np.zeros(shape),
          np.cos(X) * np.sin(Y) * np.sin(Z), np.sin(X) * np.cos(Y) * np.sin(Z), -2 * np.sin(X) * np.sin(Y) * np.cos(Z)]
    out = []
    for f in fs:
        if pinned:
            import torch
            t = torch.empty(shape, dtype=torch.float32, pin_memory=True)
            a = t.numpy()
            a[...] = f
            out.append((a, t))
        else:
            out.append((f.astype(T), None))
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
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
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


def cpu_oracle_sample(kind, n, stepper, nu, eta, dt, nsamples=1, warm=0):
    """Time the CPU oracle (literal restatement of the reference op sequence) on a bounded sample:
    one RHS evaluation (calcN!: 36/24/51 3D FFTs) of the same workload.  Returns seconds per RHS evaluation."""
    from oracle import mhdflows_oracle as O
    kw = dict(nx=n, T=np.float32, dt=dt, stepper="RK4")
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
    N = np.zeros_like(op.sol)
    times = []
    for i in range(warm + nsamples):
        t0 = time.perf_counter()
        op.calcN(N, op.sol, 0.0, op.clock, op.vars, op.params, op.grid)
        t1 = time.perf_counter()
        if i >= warm:
            times.append(t1 - t0)
        # advance a little so successive samples are not the identical input
        op.sol += np.float32(dt * 0.25) * N
    return times


def cufft_reference_point(kind, dims, reps=5):
    """cuFFT timed alongside (north_star): the 3D transforms of one RHS evaluation through torch.fft (= cuFFT plans), without
    any of the products / spectral work -- (a) the fused formulation's count (MHD 6 c2r + 9 r2c), (b) the reference's own
    count (MHD 36: 6 c2r + 30 r2c incl. the rfft(irfft()) diffusion operands).  A library reference point, never the product
    path.  Returns None when torch/cuFFT is unavailable."""
    try:
        import torch
        nx, ny, nz = dims
        n_c2r, n_r2c = {"mhd": (6, 9), "hd": (3, 6), "emhd": (24, 3)}[kind]
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
        return {"library": "cuFFT via torch.fft (out-of-place, full (N/2+1)N^2 spectra, no pruning)", "c2r_ms": c2r_ms, "r2c_ms": r2c_ms,
                "ffts_per_rhs_fused_form": n_c2r + n_r2c, "fft_only_ms_per_rhs_fused_form": fused,
                "ffts_per_rhs_reference_form": ref_total, "fft_only_ms_per_rhs_reference_form": literal,
                "what": "transforms only; our ms_per_step / stages also contains the products, the spectral assembly and the RK update"}
    except Exception as e:      # a reference point must never take the bench line down
        return {"unavailable": f"{type(e).__name__}: {e}"[:200]}


def run_reference(args, wl):
    """--impl reference: the reference's CPU path.  Julia + FFTW cannot run here (not installed, no network), so the
    oracle port (same 36-FFT op sequence, scipy pocketfft, all host threads) is timed; each "step" is a bounded
    sample = one RHS evaluation = 1/stages of a time step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    kind, n, stepper, nu, eta, dt = WORKLOADS[wl]
    stages = 4 if stepper == "RK4" else 5
    cores = os.cpu_count() or 1
    times = cpu_oracle_sample(kind, n, stepper, nu, eta, dt, nsamples=args.steps, warm=args.warmup)
    tot = float(sum(times))
    sec_per_step = tot / len(times) * stages
    val = n ** 3 / sec_per_step
    sample = f"each timed step = one calcN! evaluation ({ {'mhd': 36, 'hd': 24, 'emhd': 51}[kind]} 3D FFTs) of {wl}; a full {stepper} step = {stages} of them"
    line = {"impl": "reference", "metric": "grid-points*steps/s", "value": val, "unit": "pts*steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec_per_step * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl, "kind": kind, "n": n, "stepper": stepper,
                       "note": "reference CPU path = oracle port (NumPy + scipy.fft pocketfft); Julia/FFTW not installable offline"},
            "cpu_baseline": {"value": val, "unit": "pts*steps/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "pts*steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N > 1: weak = the workload's n^3 points per GPU (default), strong = the workload's grid split over N GPUs")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    wl = args.workload or "mhd256"
    if wl not in WORKLOADS:
        raise SystemExit(f"unknown workload {wl}; choose from {sorted(WORKLOADS)}")
    if args.impl == "reference":
        run_reference(args, wl)
        return

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if args.gpus != 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N > 1")
    import torch
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    kind, n, stepper, nu, eta, dt = WORKLOADS[wl]
    K, W = args.steps, args.warmup
    strong = args.scaling == "strong"
    dims = (n, n, n) if (world == 1 or strong) else grid_for(n, world)
    nid = None
    if world > 1:
        from mhdflows_jl_b200.dist import nccl_id_via_torch
        nid = nccl_id_via_torch()
    M, p = make_problem(kind, n, stepper, nu, eta, dt, device=local, dims=dims, rank=rank, world=world, nccl_id=nid)
    nzl = dims[2] // world
    fields = tg_fields(n, pinned=True, dims=dims, zrange=(rank * nzl, (rank + 1) * nzl))
    set_ic(M, p, kind, fields)
    S = 8 * (dims[0] // 2 + 1) * dims[1] * dims[2]      # one reference-layout spectral field of the whole grid
    npts = dims[0] * dims[1] * dims[2]

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
    x_bytes = XPASS_S_PER_LAUNCH[kind] * S // world   # this rank's share of the rows
    achieved = x_bytes / x_avg_s / 1e9
    alg_step = ALG_S_PER_STEP[(kind, stepper)] * S
    step_ach = alg_step / world / (ms_per_step * 1e-3) / 1e9   # per GPU
    shares = {k: v[0] / max(sum(x[0] for x in prof.values()), 1e-12) for k, v in prof.items() if v[1]}
    nvlink = None
    if world > 1:
        # NVLink roofline of the global transposes: contract bytes sent per GPU per step (unpruned reference-layout
        # fields) and the bytes the pruned exchange really pushes, over the exchange time measured with CUDA events on the
        # communication stream (includes the two cross-rank barriers of every exchange); peak = 900 GB/s per direction per GPU.
        try:
            stages = 4 if stepper == "RK4" else 5
            nf_x = XPASS_S_PER_LAUNCH[kind]
            contract = stages * nf_x * (S / world) * (world - 1) / world
            inf = p.info()
            real = stages * nf_x * (dims[2] // world) * p.layout.Kyl * inf["Kxp"] * 8 * (world - 1)
            ex_ms, ex_cnt = prof["exchange"]
            ex_s_per_step = ex_ms * 1e-3 / K
            nvlink = {"peak": 900.0, "unit": "GB/s", "contract_bytes_per_step_per_gpu": contract, "pushed_bytes_per_step_per_gpu": real,
                      "exchange_ms_per_step": ex_s_per_step * 1e3, "exchanges_per_step": ex_cnt / K,
                      "achieved_contract": contract / ex_s_per_step / 1e9, "frac_contract": contract / ex_s_per_step / 1e9 / 900.0,
                      "achieved_pushed": real / ex_s_per_step / 1e9, "frac_pushed": real / ex_s_per_step / 1e9 / 900.0,
                      "note": "exchange time is on the communication stream and overlaps the FFT passes; it is not additive to the step"}
        except Exception as e:
            nvlink = {"unavailable": f"{type(e).__name__}: {e}"[:200]}
    traffic = None
    tp = os.path.join(ROOT, "profiles", "xfused_traffic.json")
    if os.path.exists(tp):
        try:
            with open(tp) as f:
                traffic = json.load(f).get(wl)
        except Exception:
            traffic = None

    # ---- end-to-end through the public API with host buffers: `e2e` ----------------------
    # One timed region = SetUpProblemIC! from pinned host arrays (H2D) + K iterations of the reference's
    # TimeIntegrator! loop body (getCFL! -> stepforward! -> ProbDiagnostic, scalars D2H each step) +
    # a savefile-style download of every real field (D2H).
    nf = 3 if kind != "mhd" else 6
    p.close()
    M2, q = make_problem(kind, n, stepper, nu, eta, dt, device=local, dims=dims, rank=rank, world=world, nccl_id=nid2(world))
    set_ic(M2, q, kind, fields)
    M2.stepforward(q, 1)
    dl_bufs = []                      # pinned host buffers receiving the downloaded fields
    for _ in range(nf):
        tbuf = torch.empty(fields[0][0].shape, dtype=torch.float32, pin_memory=True)
        dl_bufs.append((tbuf.numpy(), tbuf))
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
    names = ["bx", "by", "bz"] if kind == "emhd" else (["ux", "uy", "uz", "bx", "by", "bz"][:nf])
    outs = [q.get_real(nm, M2.STALE, out=dl_bufs[i][0]) for i, nm in enumerate(names)]
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
    del outs

    if rank != 0:
        q.close()
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return
    line = {
        "metric": "grid-points*steps/s", "value": value, "unit": "pts*steps/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": (value / PUBLISHED_PTS_STEPS_PER_S[wl]) if (world == 1 and wl in PUBLISHED_PTS_STEPS_PER_S) else None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl, "kind": kind, "n": n, "grid": list(dims), "stepper": stepper, "nu": nu, "eta": eta, "dt": dt,
                   "baseline": "MHDFlows.jl README.md:78: 0.271 s/iteration, MHD TG 256^3 Float32 RK4 on an RTX 3080 (CUDA.jl)" if wl in PUBLISHED_PTS_STEPS_PER_S else None,
                   "ic": "analytic Taylor-Green u and b", "l2": "per-step working set (FFT work buffers) is far larger than the 126 MB L2; no flush needed",
                   "multi_gpu": ("slab decomposition (z slabs / ky slabs), transposes = copy-engine pushes into peer HBM over NVLink, "
                                 + ("fixed grid" if strong else f"{n}^3 points per GPU")) if world > 1 else "single"},
        "roofline": {"bound": "hbm", "kernel": "k_xfused (c2r -> products -> r2c)", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": x_bytes, "avg_launch_ms": x_avg_s * 1e3, "launches_timed": int(x_cnt),
                     "timed_in": "second pass of K steps with a CUDA event pair around every launch on the library stream",
                     "kernel_share_of_step": shares,
                     "step": {"algorithmic_bytes": alg_step, "achieved": step_ach, "frac": step_ach / peak,
                              "ms_per_step_profiled": ms_prof / K}},
        "e2e": {"value": e2e_val, "unit": "pts*steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_s / K * 1e3,
                "what": "SetUpProblemIC! from pinned host + K x (getCFL!, stepforward!, ProbDiagnostic) + download of all real fields"},
        "gpu_launches": int(l1 - l0),
        "clocks": clocks,
    }
    if nvlink is not None:
        line["nvlink"] = nvlink
    if world == 1:
        q.close()
        stages_ = 4 if stepper == "RK4" else 5
        cf = cufft_reference_point(kind, dims)
        if cf and "fft_only_ms_per_rhs_fused_form" in cf:
            cf["ours_ms_per_rhs_everything_included"] = ms_per_step / stages_
        line["cufft_ref"] = cf
    if not args.no_cpu_baseline and world == 1:
        times = cpu_oracle_sample(kind, n, stepper, nu, eta, dt, nsamples=1, warm=0)
        stages = 4 if stepper == "RK4" else 5
        sec = times[0] * stages
        line["cpu_baseline"] = {"value": npts / sec, "unit": "pts*steps/s", "cores": os.cpu_count() or 1, "kind": "port",
                                "sample": f"one calcN! evaluation of {wl} ({ {'mhd': 36, 'hd': 24, 'emhd': 51}[kind]} 3D FFTs, scipy.fft workers=all), x{stages} stages per step",
                                "ms_per_step": sec * 1e3}
    print(json.dumps(line), flush=True)
    q.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
