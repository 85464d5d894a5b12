#!/bin/bash
# strong-scaling probe: tools/dist_scale.sh NGPU N [chunk]
NG=$1; N=$2; CH=${3:-3}
MHDF_EXCH_CHUNK=$CH timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512 tools/dist_check.py $N 2>&1 | grep -E "timing|dist-vs|Error|error" 
