#!/bin/bash
# A/B of two library builds on the same box: prints ms/step + per-class times for 256^3 and 512^3 MHD, alternating.
#   tools/ab.sh [other-library]     default other library: mhdflows_jl_b200/libmhdflows_b200_prev.so
#   e.g.  python -m mhdflows_jl_b200.build --variant=f32x2 && gpurun -- 'bash tools/ab.sh mhdflows_jl_b200/libmhdflows_b200_f32x2.so'
OTHER=${1:-mhdflows_jl_b200/libmhdflows_b200_prev.so}
for rep in 1 2; do
for lib in prev new; do
  if [ $lib = prev ]; then export MHDF_LIB=$PWD/$OTHER; else unset MHDF_LIB; fi
  python - <<PY
import os, sys
sys.path.insert(0, os.getcwd())
import bench
for n in (256, 512):
    M, p = bench.make_problem("mhd", n, "RK4", 1e-3, 1e-3, 2e-4)
    bench.set_ic(M, p, "mhd", bench.tg_fields(n))
    p.step_timed(3)
    ms = p.step_timed(10) / 10
    p.profile(True); p.step_timed(10); pr = p.profile_get(); p.profile(False)
    print("$lib", n, f"{ms:.3f} ms/step |", " ".join(f"{k}={v[0]/10:.3f}" for k, v in pr.items() if v[1]), "| E", p.energy(M.FRESH))
    p.close()
PY
done
done
