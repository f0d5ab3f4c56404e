#!/bin/bash
TAG=${1:-r03f}
mkdir -p gpurun_out
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --graph 0 > gpurun_out/${TAG}_ncu_bench.log 2>&1; echo "ncu rc=$?"
