"""1024^3 MHD RK4 timing on one GPU with a cheap host-side IC (one analytic array reused for every field; timing only)."""
import os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mhdflows_jl_b200 as M
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
p = M.Problem(M.GPU(), nx=n, nu=2e-4, eta=2e-4, dt=1e-6, B_field=True)
x = (-math.pi + 2 * math.pi / n * np.arange(n)).astype(np.float32)
a = (np.sin(x).reshape(1, 1, -1) * np.cos(x).reshape(1, -1, 1) * np.cos(x).reshape(-1, 1, 1)).astype(np.float32)
for f in ("ux", "uy", "uz", "bx", "by", "bz"):
    p.set_real(f, a)
p.step_timed(2)
ms = p.step_timed(5) / 5
p.profile(True); p.step_timed(5); pr = p.profile_get(); p.profile(False)
S = 8 * (n // 2 + 1) * n * n
print(f"time mhd {n}^3 RK4: {ms:.3f} ms/step  {n**3 / ms * 1e3:.3e} pts*steps/s  contract frac={384 * S / (ms * 1e-3) / 6555.2e9:.3f}  mem={p.info()['bytes_device'] / 2**30:.1f} GiB | " +
      " ".join(f"{k}={v[0] / 5:.3f}" for k, v in pr.items() if v[1]))
