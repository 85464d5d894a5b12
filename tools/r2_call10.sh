#!/bin/bash
# Round 2, call 10 (one GPU): same-box A/B of (a) the unpadded / predicated r2c post-step of 1024-point rows against the round-1 form
# (variant r2cpad) and (b) streamed loads without L1 allocation (variant l1na), at 1024^3 and at 256^3 / 512^3.
set -u
mkdir -p gpurun_out
O=gpurun_out/r2c10
for rep in 1 2; do
  for lib in r2cpad default l1na; do
    if [ $lib = default ]; then unset MHDF_LIB; else export MHDF_LIB=$PWD/mhdflows_jl_b200/libmhdflows_b200_$lib.so; fi
    timeout 300 python tools/time1024.py 2>&1 | grep -E "^time|rror" | sed "s/^/$lib /" | tee -a ${O}_time1024.log
  done
done
unset MHDF_LIB
timeout 300 bash tools/ab.sh mhdflows_jl_b200/libmhdflows_b200_l1na.so 2>&1 | sed "s/^prev/l1na/; s/^new/default/" | tee ${O}_ab_l1na.log
MHDF_LIB=$PWD/mhdflows_jl_b200/libmhdflows_b200_l1na.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -2 | tee ${O}_pytest_l1na.log
