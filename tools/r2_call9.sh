#!/bin/bash
# Round 2, call 9 (four GPUs): the one rank count the other calls did not cover -- bit-identity, 1024^3 strong scaling, bench line
# with parity; plus the config-size digests on the final code (one GPU of the box).
set -u
mkdir -p gpurun_out
O=gpurun_out/r2c9
timeout 400 python -m pytest tests/test_gpu_configs.py -m gpu -q -rxXs > ${O}_pytest_configs.log 2>&1; tail -4 ${O}_pytest_configs.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=4 --master-addr 127.0.0.1 --master-port 29551"
timeout 200 $TR tools/dist_check.py check128 2>&1 | grep -E "dist-vs|rror" | sed "s/^/P=4 auto /" | tee ${O}_check.log
timeout 200 $TR tools/dist_time.py 1024 1024 1024 5 2>&1 | grep -E "timing|rror" | sed "s/^/1024 auto /" | tee ${O}_time.log
timeout 500 $TR bench.py --gpus 4 --steps 10 --warmup 3 > ${O}_bench4.json 2> ${O}_bench4.err; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2c9_bench4.json").read().strip().splitlines()[-1])
    print("bench N=4:", d["config"]["workload"], d["scaling"], d["ms_per_step"], "ms/step value", d["value"], "e2e", d["e2e"]["value"])
    print("parity", {k: d["parity"][k] for k in d.get("parity", {}) if k in ("max_rel_diff", "ok", "unavailable")})
    print("nvlink", d.get("nvlink"))
except Exception as e:
    print("bench parse failed", e)
PY
tail -3 ${O}_bench4.err
