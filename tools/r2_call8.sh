#!/bin/bash
# Round 2, call 8 (one GPU): suite on the final code (config digests, ND forcing, analysis, CFL), EMHD timing, racecheck / synccheck.
set -u
mkdir -p gpurun_out
O=gpurun_out/r2c8
timeout 900 python -m pytest tests -m gpu -q -rxXs > ${O}_pytest.log 2>&1; echo "pytest rc=$?" >> ${O}_pytest.log
tail -8 ${O}_pytest.log
timeout 300 python bench.py --workload emhd512 --steps 10 --warmup 3 --no-cpu-baseline > ${O}_bench_emhd512.json 2> ${O}_bench.err
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2c8_bench_emhd512.json").read().strip().splitlines()[-1])
    print("emhd512", d["ms_per_step"], {k: round(v, 3) for k, v in d["roofline"]["class_ms_per_step"].items()})
except Exception as e:
    print("bench parse failed", e)
PY
for tool in racecheck synccheck; do
  timeout 500 compute-sanitizer --tool $tool --error-exitcode 7 python tools/sanitize_target.py > ${O}_sanitize_$tool.log 2>&1
  echo "compute-sanitizer $tool rc=$?" | tee -a ${O}_sanitize_$tool.log; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize-target done" ${O}_sanitize_$tool.log | tail -3
done
