#!/bin/bash
# Round-end evidence run on ONE GPU: parity tests, smoke, bench lines (graph / eager / cfg2 / cfg5 / reference arm), ncu launch
# list, ncu --set full of the dominant kernels.  Usage: gpurun -- 'bash tools/gpu_final.sh tag'
TAG=${1:-r07z}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${TAG}_gpu.txt
timeout 700 python -m pytest tests -m gpu -q --timeout=300 > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/${TAG}_pytest.log | tail -6
timeout 200 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${TAG}_smoke.log
timeout 300 python bench.py --steps 30 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
timeout 200 python bench.py --steps 30 --warmup 5 --graph 0 --no-cpu --sustain 0 > gpurun_out/${TAG}_bench_eager.json 2>> gpurun_out/${TAG}_bench.err; echo "bench eager rc=$?"
timeout 200 python bench.py --config cfg2 --steps 30 --warmup 5 > gpurun_out/${TAG}_bench_cfg2.json 2>> gpurun_out/${TAG}_bench.err; echo "bench cfg2 rc=$?"
timeout 200 python bench.py --config cfg5 --steps 30 --warmup 5 > gpurun_out/${TAG}_bench_cfg5.json 2>> gpurun_out/${TAG}_bench.err; echo "bench cfg5 rc=$?"
timeout 200 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/${TAG}_bench_ref.json 2>/dev/null; echo "ref rc=$?"
python - <<PY
import json
for n in ("bench", "bench_eager", "bench_cfg2", "bench_cfg5", "bench_ref"):
    try:
        d = json.load(open("gpurun_out/${TAG}_%s.json" % n))
        print(n, round(d["value"]), "crops/s", round(d["ms_per_step"], 4), "ms e2e", round(d["e2e"]["value"]), "sustained", (d.get("sustained") or {}).get("value"),
              "with_adam", (d.get("with_adam") or {}).get("value"), "cpu", (d.get("cpu_baseline") or {}).get("value"), (d.get("cpu_baseline") or {}).get("kind"))
        if n == "bench":
            print("  roofline", {k: d["roofline"].get(k) for k in ("kernel", "launch_ms", "achieved", "frac", "frac_of_sustained_peak", "share_of_step")})
            print("  stages", d["roofline"]["stages_ms_per_step"])
    except Exception as e:
        print(n, "unreadable", e)
PY
tail -3 gpurun_out/${TAG}_bench.err
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --graph 0 --sustain 0"
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_bench.log 2>&1; echo "ncu launches rc=$?"
CMD="python bench.py --steps 1 --warmup 3 --no-cpu --graph 0 --sustain 0"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_conv_wgrad_kernel -s 8 -c 3 -f -o gpurun_out/${TAG}_prof_wgrad $CMD > gpurun_out/${TAG}_prof.log 2>&1; echo "wgrad rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_conv_fprop_kernel -s 20 -c 5 -f -o gpurun_out/${TAG}_prof_fprop $CMD >> gpurun_out/${TAG}_prof.log 2>&1; echo "fprop rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_ -s 12 -c 6 -f -o gpurun_out/${TAG}_prof_attn $CMD >> gpurun_out/${TAG}_prof.log 2>&1; echo "attn rc=$?"
