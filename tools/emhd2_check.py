"""The two forms of the EMHD x kernel -- MHDF_EMHD2=1 (default since round 2: multipliers in shared memory, rolled loops, 168
registers) and MHDF_EMHD2=0 (register form, the EMHD branch of k_xfused): spectral state, stale real b and CFL statistics after a
few steps must be bit-identical (same arithmetic in the same order) for RK4 / LSRK54, Float32 / Float64, 32..128-point rows.
Run under gpurun; one line per case, then timings."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mhdflows_jl_b200 as M  # noqa: E402


def run(flag, n, stepper, T, steps=3, timing=False):
    os.environ["MHDF_EMHD2"] = flag           # read when the handle is created
    p = M.Problem(M.GPU(), nx=n, ny=max(n // 2, 16), nz=2 * n if n < 128 else n, T=T, stepper=stepper, B_field=True, EMHD=True, dt=1e-4)
    rng = np.random.default_rng(11)
    shape = p._real_shape
    x = np.linspace(0, 2 * np.pi, shape[2], endpoint=False).reshape(1, 1, -1)
    y = np.linspace(0, 2 * np.pi, shape[1], endpoint=False).reshape(1, -1, 1)
    z = np.linspace(0, 2 * np.pi, shape[0], endpoint=False).reshape(-1, 1, 1)
    f = {"bx": np.sin(y + z) + 0 * x, "by": np.cos(2 * x) * np.sin(z) + 0 * y, "bz": 0.3 * np.cos(x + y) + 0 * z}
    f = {k: (v + 0.01 * rng.standard_normal(shape)).astype(T) for k, v in f.items()}
    M.SetUpProblemIC(p, **f)
    M.stepforward(p, steps)
    out = (np.array(p.sol), p.get_real("by", M.STALE), p.stale_stats()[0])
    ms = p.step_timed(5) / 5 if timing else None
    p.close()
    return out, ms


if __name__ == "__main__":
    cases = ((32, "RK4", np.float32), (64, "LSRK54", np.float32), (128, "RK4", np.float32), (32, "RK4", np.float64))
    if "--small" in sys.argv:     # dry run on the CPU-emulated library
        cases = ((16, "RK4", np.float32),)
    for n, stepper, T in cases:
        (a, _), (b, _) = run("0", n, stepper, T), run("1", n, stepper, T)
        d = max(float(np.max(np.abs(a[0] - b[0]))), float(np.max(np.abs(a[1] - b[1]))), float(np.max(np.abs(a[2] - b[2]))))
        print(f"emhd2-vs-default n={n} {stepper} {np.dtype(T).name} max abs diff {d:.3e} norm {np.linalg.norm(a[0]):.3e}", flush=True)
    if "--time" in sys.argv:
        for n in (256, 512):
            for flag in ("0", "1", "0", "1"):
                _, ms = run(flag, n, "RK4", np.float32, steps=2, timing=True)
                print(f"emhd2 timing MHDF_EMHD2={flag} {n}^3 RK4 f32: {ms:.3f} ms/step", flush=True)
