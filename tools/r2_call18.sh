#!/bin/bash
# Round 2, call 18 (two GPUs): the slab path with the 32-points-per-thread passes in their blocked addressing variants: the
# 1024^3 bench line with its parity object (2-rank run against one GPU in the same job).
set -u
mkdir -p gpurun_out
O=gpurun_out/r2c18
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29571"
timeout 300 $TR bench.py --gpus 2 --steps 6 --warmup 3 > ${O}_bench2.json 2> ${O}_bench2.err; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2c18_bench2.json").read().strip().splitlines()[-1])
    print("bench N=2:", d["config"]["workload"], d["scaling"], d["ms_per_step"], "ms/step value", d["value"])
    print("parity", {k: d["parity"][k] for k in d.get("parity", {}) if k in ("max_rel_diff", "ok", "unavailable")})
    print("classes", {k: round(v, 2) for k, v in d["roofline"]["class_ms_per_step"].items()}, "exposed", d["nvlink"]["exposed_ms_per_step"])
except Exception as e:
    print("bench parse failed", e)
PY
tail -n 2 ${O}_bench2.err
