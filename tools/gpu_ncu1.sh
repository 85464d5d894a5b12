#!/bin/bash
# ncu capture of one kernel: tools/gpu_ncu1.sh TAG WORKLOAD REGEX SKIP COUNT
set -u
mkdir -p gpurun_out
TAG=$1; WL=$2; RX=$3; SK=$4; CN=$5
ncu --set full --clock-control none --import-source on -k regex:$RX -s $SK -c $CN -f -o /tmp/cap_$TAG python tools/ncu_target.py $WL 2 > gpurun_out/ncu_$TAG.log 2>&1
ncu -i /tmp/cap_$TAG.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2>/dev/null
ncu -i /tmp/cap_$TAG.ncu-rep --page details --csv > gpurun_out/${TAG}_details.csv 2>/dev/null
ncu -i /tmp/cap_$TAG.ncu-rep --page source --csv > gpurun_out/${TAG}_source.csv 2>/dev/null
tail -3 gpurun_out/ncu_$TAG.log
