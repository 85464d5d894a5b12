#!/bin/bash
# Round 2, call 5 (two GPUs): un-joined pushes + end-only field groups -- correctness, then timings on the 8-GPU-like proxy grid
# (1024 x 1024 x 256 on 2 GPUs = the per-GPU compute of 1024^3 on 8) and on 1024^3.
set -u
mkdir -p gpurun_out
O=gpurun_out/r2c5
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29551"
MHDF_ZCHUNKS=2 MHDF_FGROUPS=2 timeout 300 $TR tools/dist_check.py check64 2>&1 | grep -E "dist-vs|rror" | sed "s/^/ZC=2 endgroups nojoin /" | tee ${O}_check.log
MHDF_ZCHUNKS=4 MHDF_FGROUPS=1 MHDF_NOJOIN=0 timeout 300 $TR tools/dist_check.py forcing64 2>&1 | grep -E "dist-vs|rror" | sed "s/^/ZC=4 groups joined /" | tee -a ${O}_check.log
run() { lab=$1; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 300 $TR tools/dist_time.py "$@" 2>&1 | grep -E "timing|rror" | sed "s/^/$lab /" | tee -a ${O}_time.log; }
run "proxy joined nogroups" MHDF_ZCHUNKS=4 MHDF_NOJOIN=0 MHDF_FGROUPS=0 -- 1024 1024 256 5
run "proxy nojoin nogroups" MHDF_ZCHUNKS=4 MHDF_FGROUPS=0 -- 1024 1024 256 5
run "proxy nojoin endgroups" MHDF_ZCHUNKS=4 MHDF_FGROUPS=2 -- 1024 1024 256 5
run "proxy nojoin groups" MHDF_ZCHUNKS=4 MHDF_FGROUPS=1 -- 1024 1024 256 5
run "proxy auto" A=1 -- 1024 1024 256 5
run "1024 nojoin nogroups" MHDF_ZCHUNKS=4 MHDF_FGROUPS=0 -- 1024 1024 1024 3
run "1024 nojoin endgroups" MHDF_ZCHUNKS=4 MHDF_FGROUPS=2 -- 1024 1024 1024 3
run "1024 auto" A=1 -- 1024 1024 1024 3
run "weak256 auto" A=1 -- 256 256 512 10
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -rxXs > ${O}_pytest_multi.log 2>&1; tail -3 ${O}_pytest_multi.log
