#!/bin/bash
# Round 2, call 12 (one GPU): direction-dependent packed passes (default) and the 16-column z-pass experiment (variant ztx16) at 1024^3.
set -u
mkdir -p gpurun_out
O=gpurun_out/r2c12
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fft or large or calcN" 2>&1 | tail -2 | tee ${O}_pytest.log
for rep in 1 2; do
  for lib in ztx16 default; do
    if [ $lib = default ]; then unset MHDF_LIB; else export MHDF_LIB=$PWD/mhdflows_jl_b200/libmhdflows_b200_$lib.so; fi
    timeout 300 python tools/time1024.py 2>&1 | grep -E "^time|rror" | sed "s/^/$lib /" | tee -a ${O}_time1024.log
  done
done
unset MHDF_LIB
MHDF_LIB=$PWD/mhdflows_jl_b200/libmhdflows_b200_ztx16.so timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fft" 2>&1 | tail -2 | tee -a ${O}_pytest.log
