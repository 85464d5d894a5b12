#!/bin/bash
# Round 2, call 3 (two GPUs): flag-synchronised, field-grouped pipeline -- correctness, then timings against the barrier form.
set -u
mkdir -p gpurun_out
O=gpurun_out/r2c3
echo "p2p: see profiles/r02_c3_p2p.log"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node=2 --master-addr 127.0.0.1 --master-port 29551"
MHDF_ZCHUNKS=2 timeout 300 $TR tools/dist_check.py check64 2>&1 | grep -E "dist-vs|rror" | sed "s/^/ZC=2 flags /" | tee ${O}_check.log
MHDF_ZCHUNKS=4 timeout 300 $TR tools/dist_check.py forcing64 2>&1 | grep -E "dist-vs|rror" | sed "s/^/ZC=4 flags /" | tee -a ${O}_check.log
t() { # label env... -- nx ny nz steps
  env "$@" 2>/dev/null
}
run() { lab=$1; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 300 $TR tools/dist_time.py "$@" 2>&1 | grep -E "timing|rror" | sed "s/^/$lab /" | tee -a ${O}_time.log; }
run "512 ZC=4 flags groups" MHDF_ZCHUNKS=4 -- 512 512 512 5
run "512 ZC=4 flags nogroups" MHDF_ZCHUNKS=4 MHDF_FGROUPS=0 -- 512 512 512 5
run "512 ZC=4 barriers nogroups" MHDF_ZCHUNKS=4 MHDF_FGROUPS=0 MHDF_FLAGS=0 -- 512 512 512 5
run "512 ZC=8 flags groups" MHDF_ZCHUNKS=8 -- 512 512 512 5
run "weak256 ZC=1 barriers" MHDF_ZCHUNKS=1 -- 256 256 512 10
run "weak256 ZC=2 flags groups" MHDF_ZCHUNKS=2 -- 256 256 512 10
run "weak256 ZC=2 flags nogroups" MHDF_ZCHUNKS=2 MHDF_FGROUPS=0 -- 256 256 512 10
run "weak256 ZC=4 flags nogroups" MHDF_ZCHUNKS=4 MHDF_FGROUPS=0 -- 256 256 512 10
run "1024 ZC=4 flags groups" MHDF_ZCHUNKS=4 -- 1024 1024 1024 3
run "1024 ZC=8 flags groups" MHDF_ZCHUNKS=8 -- 1024 1024 1024 3
MHDF_ZCHUNKS=4 timeout 600 $TR bench.py --gpus 2 --steps 5 --warmup 3 --workload mhd512 > ${O}_bench2.json 2> ${O}_bench2.err; python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2c3_bench2.json").read().strip().splitlines()[-1])
    print("bench N=2 mhd512:", d["ms_per_step"], "parity", {k: d["parity"][k] for k in d["parity"] if k in ("rel_diff", "max_rel_diff", "ok", "unavailable", "spectrum_bins_max_rel_diff")})
    print("nvlink", d.get("nvlink"))
except Exception as e:
    print("bench parse failed", e)
PY
tail -3 ${O}_bench2.err
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -rxXs > ${O}_pytest_multi.log 2>&1; tail -3 ${O}_pytest_multi.log
