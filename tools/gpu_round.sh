#!/bin/bash
# One gpurun batch: GPU tests, bench line, ncu launch list + full capture of the dominant kernels.
set -u
mkdir -p gpurun_out
TAG=${1:-r1}
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
tail -5 gpurun_out/pytest_gpu_$TAG.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; tail -c 3000 gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python tools/ncu_target.py mhd256 2 > gpurun_out/ncu_l_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_xfused -s 4 -c 1 -f -o gpurun_out/xfused_$TAG python tools/ncu_target.py mhd256 2 > gpurun_out/ncu_x_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_pass -s 20 -c 4 -f -o gpurun_out/pass_$TAG python tools/ncu_target.py mhd256 2 > gpurun_out/ncu_p_$TAG.log 2>&1
ls -la gpurun_out
