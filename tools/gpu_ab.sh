#!/bin/bash
# A/B of one library option on one GPU: parity tests with the default, then bench with the option on / off.
# Usage: gpurun -- 'bash tools/gpu_ab.sh tag "--attn-pos 1" "--attn-pos 0"'
TAG=${1:-ab}
A=${2:-}
B=${3:-}
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -q --timeout=300 > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/${TAG}_pytest.log | tail -8
grep -n "^E  " gpurun_out/${TAG}_pytest.log | head -12
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --sustain 0 $A > gpurun_out/${TAG}_bench_a.json 2> gpurun_out/${TAG}_bench.err; echo "bench A rc=$?"
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --sustain 0 $B > gpurun_out/${TAG}_bench_b.json 2>> gpurun_out/${TAG}_bench.err; echo "bench B rc=$?"
python - <<PY
import json
for n in ("bench_a", "bench_b"):
    try:
        d = json.load(open("gpurun_out/${TAG}_%s.json" % n))
        print(n, round(d["value"]), "crops/s", round(d["ms_per_step"], 4), "ms")
        print("  stages", {k: v for k, v in d["roofline"]["stages_ms_per_step"].items() if "attn" in k})
    except Exception as e:
        print(n, "unreadable", e)
PY
tail -5 gpurun_out/${TAG}_bench.err
