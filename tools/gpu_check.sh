#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench, ncu launch list.  Usage: gpurun -- 'bash tools/gpu_check.sh [tag]'
TAG=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${TAG}_gpu.txt
timeout 420 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log
timeout 240 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
timeout 240 python bench.py --steps 20 --warmup 5 --graph 0 --no-cpu > gpurun_out/${TAG}_bench_eager.json 2>> gpurun_out/${TAG}_bench.err; echo "bench eager rc=$?"
cat gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
timeout 240 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/${TAG}_bench_ref.json 2>/dev/null; cat gpurun_out/${TAG}_bench_ref.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --graph 0 > gpurun_out/${TAG}_ncu_bench.log 2>&1; echo "ncu rc=$?"
