#!/bin/bash
TAG=${1:-r03k}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu --graph 0"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:tc_conv_fprop_kernel -s 20 -c 5 -f -o gpurun_out/${TAG}_prof_fprop $CMD > gpurun_out/${TAG}_prof_fprop.log 2>&1; echo "fprop rc=$?"
