#!/bin/bash
# Parity tests, then an interleaved A/B of the graph bench under two environment settings.
# Usage: gpurun -- 'bash tools/gpu_env_ab.sh tag "VAR=1" "VAR=0" [rounds]'
TAG=${1:-env}
A=${2:-X=1}
B=${3:-X=0}
N=${4:-2}
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -q --timeout=300 > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/${TAG}_pytest.log | tail -8
grep -n "^E  " gpurun_out/${TAG}_pytest.log | head -12
for i in $(seq 1 $N); do
  for v in "$A" "$B"; do
    env $v timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu --sustain 0 > gpurun_out/${TAG}_${i}.json 2> gpurun_out/${TAG}_${i}.err
    python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_${i}.json"))
    print("$v", round(d["value"]), "crops/s", round(d["ms_per_step"], 4), "ms", "launches", d.get("gpu_launches"))
except Exception as e:
    print("$v", "unreadable", e)
PY
  done
done
