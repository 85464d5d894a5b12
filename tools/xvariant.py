"""Time the fused x kernel for one grid under the current MHDF_X_VARIANT (single GPU)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
nx, ny, nz = (int(a) for a in sys.argv[1:4])
M, p = bench.make_problem("mhd", nx, "RK4", 1e-3, 1e-3, 2e-4, dims=(nx, ny, nz))
bench.set_ic(M, p, "mhd", bench.tg_fields(nx, dims=(nx, ny, nz)))
p.step_timed(3)
ms = p.step_timed(10) / 10
p.profile(True); p.step_timed(10); pr = p.profile_get(); p.profile(False)
print(f"variant={os.environ.get('MHDF_X_VARIANT','default'):8s} {nx}x{ny}x{nz}: {ms:.3f} ms/step  x_fused={pr['x_fused'][0]/10:.3f} ms/step  E={p.energy(M.FRESH)}")
