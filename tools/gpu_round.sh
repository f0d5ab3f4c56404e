#!/bin/bash
# One GPU-box visit for a whole change set: parity tests, smoke, bench (overlap on/off, graph/eager), reference arm, ncu launch
# list.  Usage: gpurun -- 'bash tools/gpu_round.sh tag'
TAG=${1:-r03}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${TAG}_gpu.txt
timeout 540 python -m pytest tests -m gpu -q --timeout=150 > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -25 gpurun_out/${TAG}_pytest.log
timeout 200 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log
timeout 200 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
timeout 200 python bench.py --steps 20 --warmup 5 --overlap 0 --no-cpu > gpurun_out/${TAG}_bench_nooverlap.json 2>> gpurun_out/${TAG}_bench.err; echo "bench nooverlap rc=$?"
timeout 200 python bench.py --steps 20 --warmup 5 --graph 0 --no-cpu > gpurun_out/${TAG}_bench_eager.json 2>> gpurun_out/${TAG}_bench.err; echo "bench eager rc=$?"
python - <<PY
import json
for n in ("bench", "bench_nooverlap", "bench_eager"):
    try:
        d = json.load(open("gpurun_out/${TAG}_%s.json" % n))
        print(n, round(d["value"]), "crops/s", round(d["ms_per_step"], 4), "ms e2e", round(d["e2e"]["value"]), "with_adam", (d.get("with_adam") or {}).get("value"))
    except Exception as e:
        print(n, "unreadable", e)
PY
tail -5 gpurun_out/${TAG}_bench.err
timeout 200 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/${TAG}_bench_ref.json 2>/dev/null; cut -c1-300 gpurun_out/${TAG}_bench_ref.json
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --graph 0 > gpurun_out/${TAG}_ncu_bench.log 2>&1; echo "ncu rc=$?"
