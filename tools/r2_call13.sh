#!/bin/bash
# Round 2, call 13 (one GPU): 16-column z passes as the default; A/B of 16 columns in the 1024-point y passes too (variant ytx16);
# the FFT / calcN / large-grid parity tests on the new default.
set -u
mkdir -p gpurun_out
O=gpurun_out/r2c13
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fft or large or calcN" 2>&1 | tail -2 | tee ${O}_pytest.log
for rep in 1 2; do
  for lib in ytx16 default; do
    if [ $lib = default ]; then unset MHDF_LIB; else export MHDF_LIB=$PWD/mhdflows_jl_b200/libmhdflows_b200_$lib.so; fi
    timeout 300 python tools/time1024.py 2>&1 | grep -E "^time|rror" | sed "s/^/$lib /" | tee -a ${O}_time1024.log
  done
done
unset MHDF_LIB
timeout 200 python tools/time1024.py 512 2>&1 | grep -E "^time|rror" | sed "s/^/default /" | tee -a ${O}_time1024.log
