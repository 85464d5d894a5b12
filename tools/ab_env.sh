#!/bin/bash
# A/B of environment settings on the same box: tools/ab_env.sh "VAR=a" "VAR=b" ...
for rep in 1 2; do
for setting in "$@"; do
  env $setting python - <<PY
import os, sys
sys.path.insert(0, os.getcwd())
import bench
for n in (256, 512):
    M, p = bench.make_problem("mhd", n, "RK4", 1e-3, 1e-3, 2e-4)
    bench.set_ic(M, p, "mhd", bench.tg_fields(n))
    p.step_timed(3)
    ms = p.step_timed(10) / 10
    p.profile(True); p.step_timed(10); pr = p.profile_get(); p.profile(False)
    print("$setting", n, f"{ms:.3f} ms/step |", " ".join(f"{k}={v[0]/10:.3f}" for k, v in pr.items() if v[1]), "| E", p.energy(M.FRESH))
    p.close()
PY
done
done
