#!/bin/bash
# Round 2, call 11 (one GPU): strided passes on the packed float2p arithmetic (default) against the scalar forms (variant pass_scalar):
# parity suite on the new default first, then same-box A/B at 256^3 / 512^3 / 1024^3.
set -u
mkdir -p gpurun_out
O=gpurun_out/r2c11
timeout 600 python -m pytest tests -m gpu -q -x > ${O}_pytest.log 2>&1; echo "pytest rc=$?" >> ${O}_pytest.log; tail -4 ${O}_pytest.log
timeout 300 bash tools/ab.sh mhdflows_jl_b200/libmhdflows_b200_pass_scalar.so 2>&1 | sed "s/^prev/pass_scalar/; s/^new/default(packed passes)/" | tee ${O}_ab.log
for rep in 1 2; do
  for lib in pass_scalar default; do
    if [ $lib = default ]; then unset MHDF_LIB; else export MHDF_LIB=$PWD/mhdflows_jl_b200/libmhdflows_b200_$lib.so; fi
    timeout 300 python tools/time1024.py 2>&1 | grep -E "^time|rror" | sed "s/^/$lib /" | tee -a ${O}_time1024.log
  done
done
