#!/bin/bash
# Round 2, call 2 (two GPUs): slab suite incl. the z-chunk pipelined path, timings default vs pipelined, bench at N=2.
set -u
mkdir -p gpurun_out
O=gpurun_out/r2c2
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -rxXs > ${O}_pytest_multi.log 2>&1; echo "pytest rc=$?" >> ${O}_pytest_multi.log
tail -12 ${O}_pytest_multi.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29551"
for zc in 1 2 4; do
  MHDF_ZCHUNKS=$zc timeout 300 $TR tools/dist_time.py 512 512 512 5 2>&1 | grep -E "timing|Error|error" | sed "s/^/ZCHUNKS=$zc /" | tee -a ${O}_time.log
done
for zc in 1 4; do
  MHDF_ZCHUNKS=$zc timeout 400 $TR tools/dist_time.py 1024 1024 1024 3 2>&1 | grep -E "timing|Error|error" | sed "s/^/ZCHUNKS=$zc /" | tee -a ${O}_time.log
done
timeout 900 $TR bench.py --gpus 2 --steps 5 --warmup 3 > ${O}_bench2.json 2> ${O}_bench2.err; cut -c1-3000 ${O}_bench2.json; tail -5 ${O}_bench2.err
ls gpurun_out | head -40
