#!/bin/bash
# Parity tests with the in-tree library, then an interleaved A/B of prebuilt library variants (variants/*.so, loaded through
# DTA_B200_LIB; built by tools/build_variant.sh) with the graph bench.
# Usage: gpurun -- 'bash tools/gpu_variants.sh tag "v0 v1 v2 v0 v1" [notest]'
TAG=${1:-var}
ORDER=${2:-v0 v1 v0 v1}
mkdir -p gpurun_out
if [ -z "$3" ]; then
timeout 700 python -m pytest tests -m gpu -q --timeout=300 > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/${TAG}_pytest.log | tail -8
grep -n "^E  " gpurun_out/${TAG}_pytest.log | head -12
fi
i=0
for v in $ORDER; do
  i=$((i+1))
  DTA_B200_LIB=$PWD/variants/$v.so timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu --sustain 0 > gpurun_out/${TAG}_${i}_${v}.json 2> gpurun_out/${TAG}_${i}_${v}.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${TAG}_${i}_${v}.json"))
    st = d["roofline"]["stages_ms_per_step"]
    print("$v", round(d["value"]), "crops/s", round(d["ms_per_step"], 4), "ms", {k: st[k] for k in st if "conv" in k and "pack" not in k})
except Exception as e:
    print("$v", "unreadable", e)
PY
done
