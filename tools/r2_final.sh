#!/bin/bash
# Final check of the committed state (one GPU): smoke, full GPU suite, the default bench line and its reference arm.
set -u
mkdir -p gpurun_out
O=gpurun_out/r2final
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee ${O}_smoke.log
timeout 900 python -m pytest tests -m gpu -q -rxXs > ${O}_pytest.log 2>&1; echo "pytest rc=$?" >> ${O}_pytest.log; tail -7 ${O}_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > ${O}_bench.json 2> ${O}_bench.err; cut -c1-400 ${O}_bench.json; tail -2 ${O}_bench.err
