#!/bin/bash
# ncu captures; exports CSV pages on the box and drops the big .ncu-rep files (gpurun_out is capped at 64 MiB).
set -u
mkdir -p gpurun_out
TAG=${1:-r1}
WL=${2:-mhd256}
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python tools/ncu_target.py $WL 2 > gpurun_out/ncu_l_$TAG.log 2>&1
cap() { # name regex skip count
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -f -o /tmp/$1_$TAG python tools/ncu_target.py $WL 2 > gpurun_out/ncu_$1_$TAG.log 2>&1
  ncu -i /tmp/$1_$TAG.ncu-rep --page raw --csv > gpurun_out/$1_${TAG}_raw.csv 2>/dev/null
  ncu -i /tmp/$1_$TAG.ncu-rep --page details --csv > gpurun_out/$1_${TAG}_details.csv 2>/dev/null
  ncu -i /tmp/$1_$TAG.ncu-rep --page source --csv > gpurun_out/$1_${TAG}_source.csv 2>/dev/null
}
cap xfused k_xfused 4 1
cap pass k_pass 20 4
cap spectral k_spectral 4 1
ls -la gpurun_out
