#!/bin/bash
# A/B of prebuilt library variants (variants/*.so, loaded through DTA_B200_LIB) with the graph bench, interleaved.
# Usage: gpurun -- 'bash tools/gpu_variants.sh tag "v0 v2 v1 v0 v2"'
TAG=${1:-var}
ORDER=${2:-v0 v2 v1 v0 v2}
mkdir -p gpurun_out
i=0
for v in $ORDER; do
  i=$((i+1))
  DTA_B200_LIB=$PWD/variants/$v.so timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu > gpurun_out/${TAG}_${i}_${v}.json 2> gpurun_out/${TAG}_${i}_${v}.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_${i}_${v}.json"))
    print("$v", round(d["value"]), "crops/s", round(d["ms_per_step"], 4), "ms", {k: d["roofline"]["stages_ms_per_step"][k] for k in ("bwd.attn1", "bwd.attn2", "fwd.attn1", "fwd.attn2")})
except Exception as e:
    print("$v", "unreadable", e)
PY
done
