#!/bin/bash
# Round 2, call 7 (one GPU): the full GPU suite on the final code, the bench lines of every BASELINE config, the ncu launch list of
# the bench command and --set full captures of the three kernel families at the 1024^3 target.
set -u
mkdir -p gpurun_out
O=gpurun_out/r2c7
timeout 900 python -m pytest tests -m gpu -q -rxXs > ${O}_pytest.log 2>&1; echo "pytest rc=$?" >> ${O}_pytest.log
tail -8 ${O}_pytest.log
timeout 600 python bench.py --steps 20 --warmup 3 > ${O}_bench_mhd1024.json 2> ${O}_bench.err; cut -c1-700 ${O}_bench_mhd1024.json; tail -3 ${O}_bench.err
for wl in mhd256 mhd512_lsrk emhd512; do
  timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline > ${O}_bench_$wl.json 2>> ${O}_bench.err
  python - "$wl" <<'PY'
import json, sys
wl = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/r2c7_bench_{wl}.json").read().strip().splitlines()[-1])
    gb = d.get("gpu_baseline", {})
    print(wl, f"{d['ms_per_step']:.3f} ms/step value {d['value']:.4e} e2e {d['e2e']['value']:.4e} x-frac {d['roofline']['frac']:.3f} pruned step frac {d['roofline']['step']['frac_pruned']:.3f}",
          "gpu_baseline", gb.get("ms_per_step"), gb.get("ours_over_gpu_baseline"), {k: round(v, 3) for k, v in d["roofline"]["class_ms_per_step"].items()})
except Exception as e:
    print(wl, "bench parse failed", e)
PY
done
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > ${O}_bench_ref.json 2>> ${O}_bench.err; cut -c1-500 ${O}_bench_ref.json
# ncu: launch list of the bench workload (2 steps), then full captures (1024^3: one launch each; ncu replays every kernel ~40x)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ${O}_launches_mhd1024.csv python tools/ncu_target.py mhd1024 2 > ${O}_ncu_l.log 2>&1
cap() { # name workload regex skip count
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$3 -s $4 -c $5 -f -o /tmp/$1 python tools/ncu_target.py $2 2 > ${O}_ncu_$1.log 2>&1
  ncu -i /tmp/$1.ncu-rep --page raw --csv > ${O}_$1_raw.csv 2>/dev/null
  ncu -i /tmp/$1.ncu-rep --page details --csv > ${O}_$1_details.csv 2>/dev/null
  tail -2 ${O}_ncu_$1.log
}
cap xfused1024 mhd1024 k_xfused 2 1
cap pass1024 mhd1024 k_pass 12 4
cap spectral1024 mhd1024 k_spectral 2 1
cap xfused_emhd512 emhd512 k_xfused_emhd2 2 1
ls -la gpurun_out | grep r2c7 | head -40
