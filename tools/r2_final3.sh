#!/bin/bash
# Final check of the committed state after the 32-points-per-thread passes became the default (one GPU): smoke, the full GPU suite,
# the default bench line.
set -u
mkdir -p gpurun_out
O=gpurun_out/r2f3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1 | tee ${O}_smoke.log
timeout 600 python -m pytest tests -m gpu -q -rxXs > ${O}_pytest.log 2>&1; echo "pytest rc=$?" >> ${O}_pytest.log; tail -n 3 ${O}_pytest.log
timeout 400 python bench.py --steps 20 --warmup 5 > ${O}_bench_mhd1024.json 2> ${O}_bench.err; tail -n 2 ${O}_bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2f3_bench_mhd1024.json").read().strip().splitlines()[-1])
gb = d.get("gpu_baseline") or {}
print(f"mhd1024 {d['ms_per_step']:.3f} ms/step value {d['value']:.4e} e2e {d['e2e']['value']:.4e} x-frac {d['roofline']['frac']:.3f} pruned step frac {d['roofline']['step']['frac_pruned']:.3f} contract {d['roofline']['step']['contract_ratio']:.3f}",
      "gpu_baseline", gb.get("ms_per_step"), gb.get("ours_over_gpu_baseline"), {k: round(v, 3) for k, v in d["roofline"]["class_ms_per_step"].items()}, d["clocks"])
PY
