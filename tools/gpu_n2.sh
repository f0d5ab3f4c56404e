#!/bin/bash
# Two-GPU visit: parity tests (one GPU), the peer all-reduce check and the N = 2 bench.  Usage: gpurun --gpus 2 -- 'bash tools/gpu_n2.sh tag'
TAG=${1:-n2}
mkdir -p gpurun_out
timeout 540 python -m pytest tests -m gpu -q -x --timeout=150 > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest.log
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py > gpurun_out/${TAG}_dist_check.log 2>&1; echo "dist_check rc=$?"; tail -6 gpurun_out/${TAG}_dist_check.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/${TAG}_bench_n2.json 2> gpurun_out/${TAG}_bench_n2.err; echo "bench n2 rc=$?"; tail -3 gpurun_out/${TAG}_bench_n2.err
timeout 240 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; echo "bench n1 rc=$?"
python - <<PY
import json
for n in ("bench_n2", "bench_n1"):
    try:
        d = json.load(open("gpurun_out/${TAG}_%s.json" % n))
        print(n, round(d["value"]), "crops/s", round(d["ms_per_step"], 4), "ms e2e", round(d["e2e"]["value"]), d["config"].get("gradient_exchange"))
    except Exception as e:
        print(n, "unreadable", e)
PY
