"""Cost of the SURVEY 8f rows on hardware (run under gpurun): MHD 256^3 RK4 ms/step undriven vs A99-driven (both variants),
per-class times, and one DivVCorrection!/DivBCorrection! call.  Prints one line per case."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import mhdflows_jl_b200 as M  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
fields = bench.tg_fields(n)
for case in ("undriven", "a99_host", "a99_gpu"):
    kw = dict(nx=n, nu=1e-3, eta=1e-3, dt=2e-4, B_field=True)
    if case == "a99_host":
        uv, fn = M.GetA99vars_And_function(M.GPU(), n, n, n)
        kw.update(calcF=fn, usr_vars=uv)
    elif case == "a99_gpu":
        uv, fn, setup = M.A99GPU.GetA99vars_And_function(M.GPU(), n, n, n)
        kw.update(calcF=fn, usr_vars=uv)
    p = M.Problem(M.GPU(), **kw)
    bench.set_ic(M, p, "mhd", fields)
    if case == "a99_host":
        M.SetUpFk(p, kf=2, P=1e-3)
    elif case == "a99_gpu":
        setup(p, kf=2.0, P=1e-6)
    p.step_timed(3)
    ms = p.step_timed(10) / 10
    p.profile(True); p.step_timed(10); pr = p.profile_get(); p.profile(False)
    print(case, n, f"{ms:.3f} ms/step |", " ".join(f"{k}={v[0] / 10:.3f}" for k, v in pr.items() if v[1]), "| E", p.energy(M.FRESH), flush=True)
    if case == "undriven":
        t0 = time.perf_counter()
        M.DivVCorrection(p); M.DivBCorrection(p)
        print(f"DivVCorrection! + DivBCorrection! {n}^3: {(time.perf_counter() - t0) * 1e3:.3f} ms (synchronous, incl. 6 c2r for the vars refresh)", flush=True)
    p.close()
