#!/bin/bash
# Round 2, call 17 (one GPU): with 32 points per thread as the default of the 1024-point passes, A/B of (a) 16 columns in the y
# passes (ytx16), (b) packed arithmetic in the forward 1024-point passes (fwdp), (c) 32 points per thread at 512 points (e32_512,
# timed at 512^3, config-3 digest test on it).
set -u
mkdir -p gpurun_out
O=gpurun_out/r2c17
for lib in default ytx16 fwdp; do
  if [ $lib = default ]; then unset MHDF_LIB; else export MHDF_LIB=$PWD/mhdflows_jl_b200/libmhdflows_b200_$lib.so; fi
  timeout 200 python tools/time1024.py 2>&1 | grep -E "^time|rror" | sed "s/^/$lib /" | tee -a ${O}_time1024.log
done
for lib in default e32_512 default e32_512; do
  if [ $lib = default ]; then unset MHDF_LIB; else export MHDF_LIB=$PWD/mhdflows_jl_b200/libmhdflows_b200_$lib.so; fi
  timeout 100 python tools/time1024.py 512 2>&1 | grep -E "^time|rror" | sed "s/^/$lib /" | tee -a ${O}_time1024.log
done
export MHDF_LIB=$PWD/mhdflows_jl_b200/libmhdflows_b200_e32_512.so
timeout 200 python -m pytest tests/test_gpu_configs.py -m gpu -q -x -k "config3 or cfg3 or lsrk" 2>&1 | tail -n 2 | tee ${O}_pytest_e32_512.log
