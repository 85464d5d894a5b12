"""The GPU stand-in for the reference's CUDA.jl path (baseline/torch_v0.py, bench.py's `gpu_baseline`) must compute what the
CPU oracle computes -- checked here on the CPU device of torch (the same eager ops run on CUDA in the bench)."""
import numpy as np
import pytest

from baseline.torch_v0 import TorchV0
from oracle import mhdflows_oracle as O


@pytest.mark.parametrize("kind,stepper", [("mhd", "RK4"), ("hd", "RK4"), ("mhd", "LSRK54")])
def test_torch_v0_matches_the_oracle(kind, stepper):
    n, dt = 16, 5e-3
    kw = dict(nx=n, T=np.float32, nu=1e-2, dt=dt, stepper=stepper)
    op = O.Problem(eta=2e-2, B_field=True, **kw) if kind == "mhd" else O.Problem(**kw)
    ic = O.taylor_green_ic(op.grid)
    if kind == "mhd":
        O.SetUpProblemIC(op, *ic[:3], bx=ic[3], by=ic[4], bz=ic[5])
    else:
        O.SetUpProblemIC(op, *ic[:3])
    b = TorchV0(n, kind=kind, nu=1e-2, eta=2e-2, dt=dt, device="cpu", stepper=stepper)
    b.set_ic(ic)
    for _ in range(3):
        O.stepforward(op)
        b.stepforward()
    assert O.rel_l2(b.sol.numpy(), op.sol) < 1e-5
