#!/bin/bash
# Round 2, call 15 (two GPUs): the slab test-suite on the final code, including the HM89 stepper's slab check.
set -u
mkdir -p gpurun_out
O=gpurun_out/r2c15
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -rxXs > ${O}_pytest_multi.log 2>&1; echo "pytest rc=$?" >> ${O}_pytest_multi.log; tail -n 6 ${O}_pytest_multi.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29561 tools/dist_check.py hm89_64 2>&1 | grep -E "dist-vs|rror" | tee ${O}_check_hm89.log
