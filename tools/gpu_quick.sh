#!/bin/bash
# Quick GPU visit: parity tests + one graph bench.  Usage: gpurun -- 'bash tools/gpu_quick.sh tag'
TAG=${1:-q}
mkdir -p gpurun_out
timeout 540 python -m pytest tests -m gpu -q -x --timeout=150 > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/${TAG}_pytest.log
timeout 240 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_bench.json"))
print(round(d["value"]), "crops/s", round(d["ms_per_step"], 4), "ms e2e", round(d["e2e"]["value"]), "with_adam", (d.get("with_adam") or {}).get("value"))
print(d["roofline"]["stages_ms_per_step"])
PY
