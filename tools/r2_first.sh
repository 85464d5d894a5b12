#!/bin/bash
# First gpurun call of round 2: everything that was written after round 1's GPU budget was spent, in one batch.
#   (here, before the call)  python -m mhdflows_jl_b200.build --variant=f32x2 --variant=emhd_unroll
#   gpurun --timeout 2400 -- 'bash tools/r2_first.sh'
# Writes gpurun_out/r2_first_*.log.  Order: the cheap correctness checks first, then the A/B timings.
set -u
mkdir -p gpurun_out
O=gpurun_out/r2_first
# 1. full GPU suite: the xfail-guarded tests (A99 driving, divergence corrections, k_spectral2) show up as XPASS / xfail
timeout 900 python -m pytest tests -m gpu -q -rxX > ${O}_pytest.log 2>&1; echo "pytest rc=$?" >> ${O}_pytest.log
tail -15 ${O}_pytest.log
# 2. the forcing file alone with the guard lifted: real failure messages
timeout 300 python -m pytest tests/test_gpu_zforcing.py -m gpu -q --runxfail -x > ${O}_forcing.log 2>&1; echo "forcing rc=$?" >> ${O}_forcing.log
tail -5 ${O}_forcing.log
timeout 200 python tools/time_forcing.py 256 > ${O}_time_forcing.log 2>&1; cat ${O}_time_forcing.log
# 3. opt-in spectral kernel: bit-identity on hardware, then its A/B
timeout 200 python tools/spec2_check.py > ${O}_spec2_check.log 2>&1; tail -3 ${O}_spec2_check.log
timeout 300 bash tools/ab_env.sh MHDF_SPEC2=0 MHDF_SPEC2=1 > ${O}_ab_spec2.log 2>&1; cat ${O}_ab_spec2.log
# 4. packed-FP32 build (if it was built and shipped)
if [ -f mhdflows_jl_b200/libmhdflows_b200_f32x2.so ]; then
  timeout 300 bash tools/ab.sh mhdflows_jl_b200/libmhdflows_b200_f32x2.so > ${O}_ab_f32x2.log 2>&1; cat ${O}_ab_f32x2.log
  MHDF_LIB=$PWD/mhdflows_jl_b200/libmhdflows_b200_f32x2.so timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > ${O}_pytest_f32x2.log 2>&1
  tail -3 ${O}_pytest_f32x2.log
fi
# 4b. EMHD x kernel: rolled component loop (default, ~105 KB of SASS) vs the round-1 fully unrolled shape (194 KB), 256^3 and 512^3;
#     then one full ncu capture of the default EMHD x kernel (round 1 has none)
cat > /tmp/emhd_time.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import bench
for n in (256, 512):
    M, p = bench.make_problem("emhd", n, "RK4", 0.0, 0.0, 1e-5)
    bench.set_ic(M, p, "emhd", bench.tg_fields(n))
    p.step_timed(2)
    ms = p.step_timed(5) / 5
    p.profile(True); p.step_timed(5); pr = p.profile_get(); p.profile(False)
    print(os.environ.get("MHDF_LIB", "default"), "emhd", n, f"{ms:.3f} ms/step |", " ".join(f"{k}={v[0]/5:.3f}" for k, v in pr.items() if v[1]), flush=True)
    p.close()
PY
timeout 300 python /tmp/emhd_time.py > ${O}_emhd.log 2>&1
if [ -f mhdflows_jl_b200/libmhdflows_b200_emhd_unroll.so ]; then
  MHDF_LIB=$PWD/mhdflows_jl_b200/libmhdflows_b200_emhd_unroll.so timeout 300 python /tmp/emhd_time.py >> ${O}_emhd.log 2>&1
fi
cat ${O}_emhd.log
timeout 400 python tools/emhd2_check.py --time > ${O}_emhd2.log 2>&1; cat ${O}_emhd2.log
timeout 600 bash tools/gpu_ncu1.sh emhd_x_r2 emhd512 k_xfused 2 1
MHDF_EMHD2=1 timeout 600 bash tools/gpu_ncu1.sh emhd2_x_r2 emhd512 k_xfused_emhd2 2 1
# 4c. compute-sanitizer over every kernel family on tiny grids (memcheck / racecheck in the test plan)
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 7 python tools/sanitize_target.py > ${O}_sanitize_$tool.log 2>&1
  echo "compute-sanitizer $tool rc=$?" | tee -a ${O}_sanitize_$tool.log; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize-target done" ${O}_sanitize_$tool.log | tail -3
done
# 5. bench line (carries the cuFFT reference point)
timeout 400 python bench.py --steps 20 --warmup 3 > ${O}_bench.json 2> ${O}_bench.err; cut -c1-800 ${O}_bench.json; tail -3 ${O}_bench.err
ls gpurun_out | head -30
