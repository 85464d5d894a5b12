"""Timings of BASELINE configs 3 and 5 shapes on one GPU (cheap host-side IC, tiny dt: timing only)."""
import os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mhdflows_jl_b200 as M
n = 512
x = (-math.pi + 2 * math.pi / n * np.arange(n)).astype(np.float32)
a = (np.sin(x).reshape(1, 1, -1) * np.cos(x).reshape(1, -1, 1) * np.cos(x).reshape(-1, 1, 1)).astype(np.float32)
S = 8 * (n // 2 + 1) * n * n
for name, kw, fields, algS in (("MHD 512^3 LSRK54", dict(B_field=True, stepper="LSRK54", nu=5e-4, eta=5e-4), ("ux", "uy", "uz", "bx", "by", "bz"), 450),
                               ("EMHD 512^3 RK4", dict(B_field=True, EMHD=True), ("bx", "by", "bz"), 424)):
    p = M.Problem(M.GPU(), nx=n, dt=1e-7, **kw)
    for f in fields:
        p.set_real(f, a)
    p.step_timed(2)
    ms = p.step_timed(5) / 5
    p.profile(True); p.step_timed(5); pr = p.profile_get(); p.profile(False)
    print(f"time {name}: {ms:.3f} ms/step  {n**3 / ms * 1e3:.3e} pts*steps/s  contract frac={algS * S / (ms * 1e-3) / 6555.2e9:.3f}  mem={p.info()['bytes_device'] / 2**30:.1f} GiB | " +
          " ".join(f"{k}={v[0] / 5:.3f}" for k, v in pr.items() if v[1]))
    p.close()
