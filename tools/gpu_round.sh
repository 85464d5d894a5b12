#!/bin/bash
# One gpurun batch: GPU tests, bench line, ncu launch list + full captures exported as CSV (gpurun_out is capped at 64 MiB).
set -u
mkdir -p gpurun_out
TAG=${1:-r1}
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
tail -4 gpurun_out/pytest_gpu_$TAG.log
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; cut -c1-600 gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err; cut -c1-400 gpurun_out/bench_ref_$TAG.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$TAG.csv python tools/ncu_target.py mhd256 2 > gpurun_out/ncu_l_$TAG.log 2>&1
cap() { # name regex skip count
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -f -o /tmp/$1_$TAG python tools/ncu_target.py mhd256 2 > gpurun_out/ncu_$1_$TAG.log 2>&1
  ncu -i /tmp/$1_$TAG.ncu-rep --page raw --csv > gpurun_out/$1_${TAG}_raw.csv 2>/dev/null
  ncu -i /tmp/$1_$TAG.ncu-rep --page details --csv > gpurun_out/$1_${TAG}_details.csv 2>/dev/null
  ncu -i /tmp/$1_$TAG.ncu-rep --page source --csv > gpurun_out/$1_${TAG}_source.csv 2>/dev/null
}
cap xfused k_xfused 4 1
cap pass k_pass 20 4
cap spectral k_spectral 4 1
ls gpurun_out | head -30
