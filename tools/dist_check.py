"""Slab-decomposed run vs the single-GPU run of the same problem (launch with torchrun, one rank per GPU).
Every rank builds the P-rank problem; rank 0 also builds the 1-GPU problem and compares gathered results."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch, torch.distributed as dist
import mhdflows_jl_b200 as M
from mhdflows_jl_b200.dist import SlabLayout, nccl_id_via_torch
from oracle import mhdflows_oracle as O

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
r, P = dist.get_rank(), dist.get_world_size()


def run(kind, n, stepper, T, nsteps, timing=False, driven=False):
    kw = dict(nx=n, T=T, stepper=stepper, dt=2e-3)
    uvs = []
    if driven:      # A99 random driving: the Philox counter is the global mode index, so slabs draw the single-GPU numbers
        for _ in range(2):
            uv, fn = M.GetA99vars_And_function(M.GPU(local), n, n, n, T=T, seed=77)
            uvs.append(uv)
        kw.update(calcF=fn)
    if kind == "mhd":
        kw.update(nu=2e-2, eta=3e-2, B_field=True)
    elif kind == "hd":
        kw.update(nu=2e-2)
    else:
        kw.update(B_field=True, EMHD=True, dt=1e-4)
    g = O.Grid(n, T=T)
    u, b = O.random_phase_ic(g, 1234), O.random_phase_ic(g, 5678)
    ic = dict(bx=b[0], by=b[1], bz=b[2]) if kind == "emhd" else (dict(ux=u[0], uy=u[1], uz=u[2], bx=b[0], by=b[1], bz=b[2]) if kind == "mhd" else dict(ux=u[0], uy=u[1], uz=u[2]))
    nid = nccl_id_via_torch()
    dp = M.Problem(M.GPU(local), rank=r, nranks=P, nccl_id=nid, **(dict(kw, usr_vars=uvs[0]) if driven else kw))
    lay = dp.layout
    if os.environ.get("MHDF_PEER", "1") == "1":
        from mhdflows_jl_b200.dist import enable_peer_exchange
        enable_peer_exchange(dp)
    M.SetUpProblemIC(dp, **{k: lay.scatter_real(v) for k, v in ic.items()})
    if driven:
        M.SetUpFk(dp, kf=3, P=2)
        M.DivVCorrection(dp)
        M.DivBCorrection(dp)
    M.stepforward(dp, nsteps)
    slabs = [dp.get_spectral(i) for i in range(dp.Nl)]
    reals = dp.get_real(0, M.STALE)
    en = dp.energy(M.STALE), dp.energy(M.FRESH), dp.helicity(), dp.stale_stats()
    gathered = [None] * P
    dist.all_gather_object(gathered, (slabs, reals))
    if r == 0:
        sp = M.Problem(M.GPU(local), **(dict(kw, usr_vars=uvs[1]) if driven else kw))
        M.SetUpProblemIC(sp, **ic)
        if driven:
            M.SetUpFk(sp, kf=3, P=2)
            M.DivVCorrection(sp)
            M.DivBCorrection(sp)
        M.stepforward(sp, nsteps)
        worst = 0.0
        for i in range(sp.Nl):
            full = lay.assemble_spectral([gth[0][i] for gth in gathered])
            ref = sp.get_spectral(i)
            worst = max(worst, float(np.abs(full - ref).max() / np.abs(ref).max()))
        re = np.concatenate([gth[1] for gth in gathered], axis=0)
        rref = sp.get_real(0, M.STALE)
        rerr = float(np.abs(re - rref).max() / np.abs(rref).max())
        e1 = sp.energy(M.STALE), sp.energy(M.FRESH), sp.helicity(), sp.stale_stats()
        ediff = max(abs(a - b) / (abs(b) + 1e-30) for x, y in zip(en[:3], e1[:3]) for a, b in zip(x, y))
        sdiff = float(max(np.abs(en[3][0] - e1[3][0]).max(), 0))
        print(f"dist-vs-single {'driven ' if driven else ''}{kind} {n}^3 {stepper} {np.dtype(T).name} P={P}: spectral max rel diff {worst:.2e}  real {rerr:.2e}  diag rel {ediff:.2e}  maxsq abs {sdiff:.2e}", flush=True)
        sp.close()
    if timing:
        dp.step_timed(2)
        dist.barrier(); torch.cuda.synchronize()
        ms = dp.step_timed(5) / 5
        t = torch.tensor([ms], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dp.profile(True); dp.step_timed(5); pr = dp.profile_get(); dp.profile(False)
        if r == 0:
            print(f"   timing P={P} {kind} {n}^3: {t.item():.3f} ms/step  {n**3 / t.item() * 1e3:.3e} pts*steps/s ; per class ms/step: " +
                  " ".join(f"{k}={v[0] / 5:.3f}" for k, v in pr.items() if v[1]), flush=True)
    dp.close()
    dist.barrier()


which = sys.argv[1] if len(sys.argv) > 1 else "check"
if which == "check64":
    run("mhd", 64, "RK4", np.float32, 3)
    run("hd", 64, "LSRK54", np.float32, 2)
    run("emhd", 64, "RK4", np.float64, 2)
elif which == "check128":      # smallest cubic grid 8 ranks can split (every rank needs >= 2 retained ky rows and a non-empty last slab)
    run("mhd", 128, "RK4", np.float32, 3)
    run("emhd", 128, "LSRK54", np.float32, 2)
elif which == "hm89_64":       # the EMHD HM89 stepper: its fixed-point error norm is a max over the ranks; iteration counts must agree
    run("emhd", 64, "HM89", np.float32, 2)
    run("emhd", 64, "HM89", np.float64, 2)
elif which == "forcing64":
    run("mhd", 64, "RK4", np.float32, 3, driven=True)
    run("mhd", 64, "LSRK54", np.float64, 2, driven=True)
elif which == "check":
    run("mhd", 64, "RK4", np.float32, 3)
    run("hd", 64, "LSRK54", np.float32, 2)
    run("emhd", 64, "RK4", np.float64, 2)
    run("mhd", 128, "RK4", np.float32, 2, timing=True)
    run("mhd", 256, "RK4", np.float32, 2, timing=True)
else:
    n = int(which)
    run("mhd", n, "RK4", np.float32, 2, timing=True)
dist.destroy_process_group()
