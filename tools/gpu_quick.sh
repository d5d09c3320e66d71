#!/bin/bash
# quick GPU check: parity tests + bench without the CPU baseline + per-CTA cycle breakdown
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench.json'))
print("value", round(d["value"],1), "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],1), "roofline", round(d["roofline"]["frac"],3), d["roofline"]["families_ms_per_step"])
PY
timeout 120 python tools/cta_timing.py > gpurun_out/cta_timing.txt 2>&1; cat gpurun_out/cta_timing.txt
