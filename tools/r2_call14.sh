#!/bin/bash
# Round 2, call 14 (two GPUs): the 16-column 1024-point z passes in their blocked (slab) addressing variants and the packed passes:
# bit-identity on small grids, then the 1024^3 bench line with its parity object (N-rank run vs one GPU) and the strong-scaling time.
set -u
mkdir -p gpurun_out
O=gpurun_out/r2c14
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29551"
timeout 300 $TR tools/dist_check.py check64 2>&1 | grep -E "dist-vs|rror" | tee ${O}_check.log
timeout 600 $TR bench.py --gpus 2 --steps 8 --warmup 3 > ${O}_bench2.json 2> ${O}_bench2.err; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2c14_bench2.json").read().strip().splitlines()[-1])
    print("bench N=2:", d["config"]["workload"], d["scaling"], d["ms_per_step"], "ms/step value", d["value"])
    print("parity", {k: d["parity"][k] for k in d.get("parity", {}) if k in ("max_rel_diff", "ok", "unavailable")})
    print("classes", {k: round(v, 2) for k, v in d["roofline"]["class_ms_per_step"].items()})
except Exception as e:
    print("bench parse failed", e)
PY
tail -3 ${O}_bench2.err
