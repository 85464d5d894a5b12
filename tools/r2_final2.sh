#!/bin/bash
# Final check of the committed state (one GPU): smoke, the full GPU suite (incl. the HM89 stepper's first hardware run), the
# default bench line, the bench lines of the other BASELINE configs, the reference arm, the ncu launch list of the bench workload and
# a --set full capture of the re-shaped 1024-point strided passes, HM89 timing.
set -u
mkdir -p gpurun_out
O=gpurun_out/r2f2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2 | tee ${O}_smoke.log
timeout 900 python -m pytest tests -m gpu -q -rxXs > ${O}_pytest.log 2>&1; echo "pytest rc=$?" >> ${O}_pytest.log; tail -n 8 ${O}_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > ${O}_bench_mhd1024.json 2> ${O}_bench.err; cut -c1-400 ${O}_bench_mhd1024.json; tail -n 2 ${O}_bench.err
for wl in mhd256 mhd512_lsrk emhd512; do
  timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline > ${O}_bench_$wl.json 2>> ${O}_bench.err
done
python - <<'PY'
import json
for wl in ("mhd1024", "mhd256", "mhd512_lsrk", "emhd512"):
    try:
        d = json.loads(open(f"gpurun_out/r2f2_bench_{wl}.json").read().strip().splitlines()[-1])
        gb = d.get("gpu_baseline") or {}
        print(wl, f"{d['ms_per_step']:.3f} ms/step value {d['value']:.4e} e2e {d['e2e']['value']:.4e} x-frac {d['roofline']['frac']:.3f} pruned step frac {d['roofline']['step']['frac_pruned']:.3f} contract {d['roofline']['step']['contract_ratio']:.3f}",
              "gpu_baseline", gb.get("ms_per_step"), gb.get("ours_over_gpu_baseline"), {k: round(v, 3) for k, v in d["roofline"]["class_ms_per_step"].items()})
    except Exception as e:
        print(wl, "bench parse failed", e)
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > ${O}_bench_ref.json 2>> ${O}_bench.err; cut -c1-300 ${O}_bench_ref.json
timeout 200 python tools/hm89_time.py 256 5 2>&1 | tail -n 3 | tee ${O}_hm89_time.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${O}_launches_mhd1024.csv python tools/ncu_target.py mhd1024 2 > ${O}_ncu_l.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_pass -s 12 -c 4 -f -o /tmp/pass1024 python tools/ncu_target.py mhd1024 2 > ${O}_ncu_pass.log 2>&1
ncu -i /tmp/pass1024.ncu-rep --page raw --csv > ${O}_pass1024_raw.csv 2>/dev/null
ncu -i /tmp/pass1024.ncu-rep --page details --csv > ${O}_pass1024_details.csv 2>/dev/null
tail -n 2 ${O}_ncu_pass.log
ls -la gpurun_out | grep r2f2
