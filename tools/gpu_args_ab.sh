#!/bin/bash
# Parity tests (optionally a -k filter first), then an interleaved A/B of the graph bench under two argument sets.
# Usage: gpurun -- 'bash tools/gpu_args_ab.sh tag "--step fused" "--step autograd" [rounds] [pytest -k filter]'
TAG=${1:-args}
A=${2:-}
B=${3:-}
N=${4:-2}
K=${5:-}
mkdir -p gpurun_out
if [ -n "$K" ]; then timeout 300 python -m pytest tests -m gpu -q -x -k "$K" --timeout=120 2>&1 | tail -15; fi
timeout 700 python -m pytest tests -m gpu -q --timeout=300 > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/${TAG}_pytest.log | tail -8
grep -n "^E  " gpurun_out/${TAG}_pytest.log | head -12
for i in $(seq 1 $N); do
  for v in "$A" "$B"; do
    if [ "$v" == "$A" ]; then i=${i%[ab]}a; else i=${i%[ab]}b; fi
    timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu --sustain 0 $v > gpurun_out/${TAG}_${i}.json 2> gpurun_out/${TAG}_${i}.err
    python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_${i}.json"))
    print("$v", round(d["value"]), "crops/s", round(d["ms_per_step"], 4), "ms", "launches", d.get("gpu_launches"), "e2e", round(d["e2e"]["value"]), "adam", (d.get("with_adam") or {}).get("value"))
except Exception as e:
    print("$v", "unreadable", e)
    print(open("gpurun_out/${TAG}_${i}.err").read()[-1500:])
PY
  done
done
