"""Short target for ncu: N^3 problem, a few steps (kept tiny because ncu replays every kernel ~40x)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench

wl = sys.argv[1] if len(sys.argv) > 1 else "mhd256"
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
kind, n, stepper, nu, eta, dt = bench.WORKLOADS[wl]
M, p = bench.make_problem(kind, n, stepper, nu, eta, dt)
bench.set_ic(M, p, kind, bench.tg_fields(n))
M.stepforward(p, nsteps)
print("done", p.energy(M.FRESH))
