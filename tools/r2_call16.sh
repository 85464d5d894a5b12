#!/bin/bash
# Round 2, call 16 (one GPU): A/B of 32 points per thread in the 1024-point strided passes (radix 32 x 32, one shared-memory
# exchange instead of two): variant e32y = the y passes, e32yz = y and z passes; FFT / calcN / large-grid parity on e32yz.
set -u
mkdir -p gpurun_out
O=gpurun_out/r2c16
for lib in default e32y e32yz; do
  if [ $lib = default ]; then unset MHDF_LIB; else export MHDF_LIB=$PWD/mhdflows_jl_b200/libmhdflows_b200_$lib.so; fi
  timeout 200 python tools/time1024.py 2>&1 | grep -E "^time|rror" | sed "s/^/$lib /" | tee -a ${O}_time1024.log
done
export MHDF_LIB=$PWD/mhdflows_jl_b200/libmhdflows_b200_e32yz.so
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fft or large or calcN" 2>&1 | tail -n 2 | tee ${O}_pytest_e32yz.log
