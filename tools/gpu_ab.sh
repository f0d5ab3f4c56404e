#!/bin/bash
# A/B of a library option on the GPU box: tests, then the graph bench with the option on and off.
# Usage: gpurun -- 'bash tools/gpu_ab.sh tag "--pdl 0"'
TAG=${1:-ab}
ALT=${2:---pdl 0}
mkdir -p gpurun_out
timeout 540 python -m pytest tests -m gpu -q -x --timeout=150 > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest.log
timeout 240 python bench.py --steps 30 --warmup 5 --no-cpu > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/${TAG}_bench.err
timeout 240 python bench.py --steps 30 --warmup 5 --no-cpu $ALT > gpurun_out/${TAG}_bench_alt.json 2>> gpurun_out/${TAG}_bench.err; echo "bench alt rc=$?"
python - <<PY
import json
for n in ("bench", "bench_alt"):
    d = json.load(open("gpurun_out/${TAG}_%s.json" % n))
    print(n, round(d["value"]), "crops/s", round(d["ms_per_step"], 4), "ms e2e", round(d["e2e"]["value"]), "raw", round(d["e2e_raw_int16"]["value"]))
PY
