#!/bin/bash
# Round 2, call 4 (eight GPUs, charged 8-fold: keep it short): correctness at P=8, 1024^3 strong-scaling variants, bench line.
set -u
mkdir -p gpurun_out
O=gpurun_out/r2c4
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=8 --master-addr 127.0.0.1 --master-port 29551"
timeout 200 $TR tools/dist_check.py check64 2>&1 | grep -E "dist-vs|rror" | sed "s/^/P=8 auto /" | tee ${O}_check.log
run() { lab=$1; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 200 $TR tools/dist_time.py "$@" 2>&1 | grep -E "timing|rror" | sed "s/^/$lab /" | tee -a ${O}_time.log; }
run "1024 auto" A=1 -- 1024 1024 1024 4
run "1024 ZC=4 nogroups" MHDF_ZCHUNKS=4 MHDF_FGROUPS=0 -- 1024 1024 1024 4
run "1024 ZC=2 groups" MHDF_ZCHUNKS=2 MHDF_FGROUPS=1 -- 1024 1024 1024 4
run "1024 ZC=8 nogroups" MHDF_ZCHUNKS=8 MHDF_FGROUPS=0 -- 1024 1024 1024 4
run "1024 ZC=4 nogroups barriers" MHDF_ZCHUNKS=4 MHDF_FGROUPS=0 MHDF_FLAGS=0 -- 1024 1024 1024 4
run "weak512 auto" A=1 -- 512 512 512 10
run "weak512 ZC=1 barriers" MHDF_ZCHUNKS=1 -- 512 512 512 10
run "weak512 ZC=4 nogroups" MHDF_ZCHUNKS=4 MHDF_FGROUPS=0 -- 512 512 512 10
timeout 500 $TR bench.py --gpus 8 --steps 10 --warmup 3 > ${O}_bench8.json 2> ${O}_bench8.err; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2c4_bench8.json").read().strip().splitlines()[-1])
    print("bench N=8:", d["config"]["workload"], d["ms_per_step"], "ms/step value", d["value"], "e2e", d["e2e"]["value"])
    print("parity", {k: d["parity"][k] for k in d.get("parity", {}) if k in ("max_rel_diff", "ok", "unavailable")})
    print("nvlink", d.get("nvlink"))
    print("classes", d["roofline"]["class_ms_per_step"])
except Exception as e:
    print("bench parse failed", e)
PY
tail -3 ${O}_bench8.err
