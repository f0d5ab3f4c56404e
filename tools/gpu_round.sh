#!/bin/bash
# One GPU-box visit for a whole change set: parity tests, smoke, bench (graph / eager), the other BASELINE configs, reference arm,
# ncu launch list.  Usage: gpurun -- 'bash tools/gpu_round.sh tag [quick]'
TAG=${1:-r04}
QUICK=${2:-}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${TAG}_gpu.txt
timeout 700 python -m pytest tests -m gpu -q --timeout=300 > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/${TAG}_pytest.log | tail -15
timeout 200 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${TAG}_smoke.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
if [ -z "$QUICK" ]; then
timeout 200 python bench.py --steps 20 --warmup 5 --graph 0 --no-cpu --sustain 0 > gpurun_out/${TAG}_bench_eager.json 2>> gpurun_out/${TAG}_bench.err; echo "bench eager rc=$?"
timeout 200 python bench.py --config cfg2 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_cfg2.json 2>> gpurun_out/${TAG}_bench.err; echo "bench cfg2 rc=$?"
timeout 200 python bench.py --config cfg5 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_cfg5.json 2>> gpurun_out/${TAG}_bench.err; echo "bench cfg5 rc=$?"
timeout 200 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/${TAG}_bench_ref.json 2>/dev/null; cut -c1-200 gpurun_out/${TAG}_bench_ref.json
fi
python - <<PY
import json
for n in ("bench", "bench_eager", "bench_cfg2", "bench_cfg5"):
    try:
        d = json.load(open("gpurun_out/${TAG}_%s.json" % n))
        print(n, round(d["value"]), "crops/s", round(d["ms_per_step"], 4), "ms e2e", round(d["e2e"]["value"]), "sustained", (d.get("sustained") or {}).get("value"),
              "with_adam", (d.get("with_adam") or {}).get("value"), "cpu", (d.get("cpu_baseline") or {}).get("value"), (d.get("cpu_baseline") or {}).get("kind"))
        if n == "bench":
            print("  stages", d["roofline"]["stages_ms_per_step"])
    except Exception as e:
        print(n, "unreadable", e)
PY
tail -5 gpurun_out/${TAG}_bench.err
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --graph 0 --sustain 0 > gpurun_out/${TAG}_ncu_bench.log 2>&1; echo "ncu rc=$?"
