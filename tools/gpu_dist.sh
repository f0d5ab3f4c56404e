#!/bin/bash
# Multi-GPU check: gradient exchange correctness (tools/dist_check.py) and the bench line at N GPUs with the exchange as one kernel
# after the backward pass (default) and, for comparison, inside it.  Usage: gpurun --gpus N -- 'bash tools/gpu_dist.sh tag N'
TAG=${1:-r05}
N=${2:-2}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $RUN --master-port 29511 tools/dist_check.py > gpurun_out/${TAG}_dist_check_n${N}.log 2>&1; grep -E "paths|exchange|graph replay|fused|DIST CHECK|Error|error" gpurun_out/${TAG}_dist_check_n${N}.log | tail -8
timeout 300 $RUN --master-port 29512 bench.py --gpus $N --steps 30 --warmup 5 --sustain 0 > gpurun_out/${TAG}_bench_n${N}.json 2> gpurun_out/${TAG}_bench_n${N}.err; echo "bench rc=$?"
DTA_EXCHANGE_IN_BACKWARD=1 timeout 300 $RUN --master-port 29513 bench.py --gpus $N --steps 30 --warmup 5 --sustain 0 > gpurun_out/${TAG}_bench_n${N}_inbackward.json 2>> gpurun_out/${TAG}_bench_n${N}.err; echo "bench(exchange in backward) rc=$?"
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu --sustain 0 > gpurun_out/${TAG}_bench_n1.json 2>> gpurun_out/${TAG}_bench_n${N}.err
python - <<PY
import json
base = None
for n in ("bench_n1", "bench_n${N}", "bench_n${N}_inbackward"):
    try:
        d = json.load(open("gpurun_out/${TAG}_%s.json" % n))
        if base is None:
            base = d["value"]
        print(n, round(d["value"]), "crops/s", round(d["ms_per_step"], 4), "ms", "eff", round(d["value"] / (base * d["n_gpus"]), 4), "e2e", round(d["e2e"]["value"]),
              d["config"].get("gradient_exchange"), d.get("grad_sync_check"))
    except Exception as e:
        print(n, "unreadable", e)
PY
tail -3 gpurun_out/${TAG}_bench_n${N}.err
