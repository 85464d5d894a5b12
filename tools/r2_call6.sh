#!/bin/bash
# Round 2, call 6 (eight GPUs, charged 8-fold): final slab pipeline -- bit-identity at P=8, 1024^3 strong scaling, weak 256^3/GPU.
set -u
mkdir -p gpurun_out
O=gpurun_out/r2c6
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=8 --master-addr 127.0.0.1 --master-port 29551"
timeout 200 $TR tools/dist_check.py check128 2>&1 | grep -E "dist-vs|rror" | sed "s/^/P=8 auto /" | tee ${O}_check.log
run() { lab=$1; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 200 $TR tools/dist_time.py "$@" 2>&1 | grep -E "timing|rror" | sed "s/^/$lab /" | tee -a ${O}_time.log; }
run "1024 auto (nojoin, end groups)" A=1 -- 1024 1024 1024 5
run "1024 nojoin nogroups" MHDF_FGROUPS=0 -- 1024 1024 1024 5
run "1024 joined nogroups (call-4 best)" MHDF_NOJOIN=0 MHDF_FGROUPS=0 -- 1024 1024 1024 5
run "weak512 auto" A=1 -- 512 512 512 10
