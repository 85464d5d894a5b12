#!/bin/bash
# Round 2, call 1 (one GPU): suite with -rxX, k_spectral2 check + A/B, packed-FP32 A/B, EMHD kernel shapes, 1024^3 timing.
set -u
mkdir -p gpurun_out
O=gpurun_out/r2c1
timeout 600 python -m pytest tests -m gpu -q -rxX > ${O}_pytest.log 2>&1; echo "pytest rc=$?" >> ${O}_pytest.log
tail -25 ${O}_pytest.log
timeout 200 python tools/spec2_check.py > ${O}_spec2_check.log 2>&1; cat ${O}_spec2_check.log | tail -8
timeout 300 bash tools/ab_env.sh MHDF_SPEC2=0 MHDF_SPEC2=1 > ${O}_ab_spec2.log 2>&1; cat ${O}_ab_spec2.log
if [ -f mhdflows_jl_b200/libmhdflows_b200_f32x2.so ]; then
  timeout 300 bash tools/ab.sh mhdflows_jl_b200/libmhdflows_b200_f32x2.so > ${O}_ab_f32x2.log 2>&1; cat ${O}_ab_f32x2.log
  MHDF_LIB=$PWD/mhdflows_jl_b200/libmhdflows_b200_f32x2.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > ${O}_pytest_f32x2.log 2>&1
  tail -3 ${O}_pytest_f32x2.log
fi
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
    print(os.path.basename(os.environ.get("MHDF_LIB", "default")), "EMHD2=" + os.environ.get("MHDF_EMHD2", "0"), "emhd", n, f"{ms:.3f} ms/step |", " ".join(f"{k}={v[0]/5:.3f}" for k, v in pr.items() if v[1]), flush=True)
    p.close()
PY
timeout 200 python /tmp/emhd_time.py > ${O}_emhd.log 2>&1
MHDF_EMHD2=1 timeout 200 python /tmp/emhd_time.py >> ${O}_emhd.log 2>&1
if [ -f mhdflows_jl_b200/libmhdflows_b200_emhd_unroll.so ]; then
  MHDF_LIB=$PWD/mhdflows_jl_b200/libmhdflows_b200_emhd_unroll.so timeout 200 python /tmp/emhd_time.py >> ${O}_emhd.log 2>&1
fi
if [ -f mhdflows_jl_b200/libmhdflows_b200_f32x2.so ]; then
  MHDF_LIB=$PWD/mhdflows_jl_b200/libmhdflows_b200_f32x2.so timeout 200 python /tmp/emhd_time.py >> ${O}_emhd.log 2>&1
fi
cat ${O}_emhd.log
timeout 300 python tools/time1024.py > ${O}_time1024.log 2>&1; cat ${O}_time1024.log
MHDF_SPEC2=1 timeout 300 python tools/time1024.py > ${O}_time1024_spec2.log 2>&1; cat ${O}_time1024_spec2.log
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_target.py > ${O}_sanitize_memcheck.log 2>&1
echo "compute-sanitizer memcheck rc=$?" | tee -a ${O}_sanitize_memcheck.log; grep -E "ERROR SUMMARY|sanitize-target done" ${O}_sanitize_memcheck.log | tail -3
ls gpurun_out | head -40
