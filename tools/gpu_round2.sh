#!/bin/bash
# GPU-box visit: parity tests, graph bench, ncu --set full captures of the attention kernels (source-level).
# Usage: gpurun -- 'bash tools/gpu_round2.sh tag'
TAG=${1:-r03b}
mkdir -p gpurun_out
timeout 540 python -m pytest tests -m gpu -q --timeout=150 > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -15 gpurun_out/${TAG}_pytest.log
timeout 200 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_bench.json"))
print(round(d["value"]), "crops/s", round(d["ms_per_step"], 4), "ms e2e", round(d["e2e"]["value"]), "with_adam", (d.get("with_adam") or {}).get("value"))
print(d["roofline"]["stages_ms_per_step"])
PY
tail -3 gpurun_out/${TAG}_bench.err
CMD="python bench.py --steps 1 --warmup 3 --no-cpu --graph 0"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_kernel -s 9 -c 3 -f -o gpurun_out/${TAG}_prof_attn_bwd $CMD > gpurun_out/${TAG}_prof_attn.log 2>&1; echo "attn_bwd rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_kernel -s 9 -c 3 -f -o gpurun_out/${TAG}_prof_attn_fwd $CMD >> gpurun_out/${TAG}_prof_attn.log 2>&1; echo "attn_fwd rc=$?"
ls -la gpurun_out/*.ncu-rep
