"""Cheap multi-GPU timing probe (torchrun): analytic Taylor-Green slabs, no oracle, no big host arrays.
usage: dist_time.py NX NY NZ [steps]"""
import os, sys, time, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch, torch.distributed as dist
import mhdflows_jl_b200 as M
from mhdflows_jl_b200.dist import nccl_id_via_torch

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
r, P = dist.get_rank(), dist.get_world_size()
nx, ny, nz = (int(a) for a in sys.argv[1:4])
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
kw = dict(nx=nx, ny=ny, nz=nz, Lx=2 * math.pi, Ly=2 * math.pi * ny / nx, Lz=2 * math.pi * nz / nx, nu=1e-3, eta=1e-3, dt=2e-4, B_field=True)
if P > 1:
    p = M.Problem(M.GPU(local), rank=r, nranks=P, nccl_id=nccl_id_via_torch(), **kw)
    nzl = p.layout.nzl
    if os.environ.get("MHDF_PEER", "1") == "1":
        from mhdflows_jl_b200.dist import enable_peer_exchange
        enable_peer_exchange(p)
else:
    p = M.Problem(M.GPU(local), **kw)
    nzl = nz
x = (-math.pi + 2 * math.pi / nx * np.arange(nx)).reshape(1, 1, -1)
y = (-math.pi * ny / nx + 2 * math.pi / nx * np.arange(ny)).reshape(1, -1, 1)
z = (-math.pi * nz / nx + 2 * math.pi / nx * np.arange(r * nzl, (r + 1) * nzl)).reshape(-1, 1, 1)
f32 = np.float32
M.SetUpProblemIC(p, ux=(np.sin(x) * np.cos(y) * np.cos(z)).astype(f32), uy=(-np.cos(x) * np.sin(y) * np.cos(z)).astype(f32),
                 uz=np.zeros((nzl, ny, nx), f32), bx=(np.cos(x) * np.sin(y) * np.sin(z)).astype(f32),
                 by=(np.sin(x) * np.cos(y) * np.sin(z)).astype(f32), bz=(-2 * np.sin(x) * np.sin(y) * np.cos(z)).astype(f32))
p.step_timed(2)
dist.barrier(); torch.cuda.synchronize()
ms = p.step_timed(steps) / steps
t = torch.tensor([ms], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
p.profile(True); p.step_timed(steps); pr = p.profile_get(); p.profile(False)
e = p.energy(M.FRESH)
if r == 0:
    print(f"timing P={P} mhd {nx}x{ny}x{nz} chunk={os.environ.get('MHDF_EXCH_CHUNK','3')} peer={os.environ.get('MHDF_PEER','1')} ncs={os.environ.get('MHDF_COPY_STREAMS','7')}: {t.item():.3f} ms/step  {nx*ny*nz / t.item() * 1e3:.3e} pts*steps/s  E={e[0]:.4f},{e[1]:.4f} ; ms/step: " +
          " ".join(f"{k}={v[0] / steps:.3f}" for k, v in pr.items() if v[1]), flush=True)
p.close()
dist.destroy_process_group()
