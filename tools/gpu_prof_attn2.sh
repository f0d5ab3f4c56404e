#!/bin/bash
TAG=${1:-r03d}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu --graph 0"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_kernel -s 9 -c 3 -f -o gpurun_out/${TAG}_prof_attn_bwd $CMD > gpurun_out/${TAG}_prof_attn.log 2>&1; echo "attn_bwd rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_kernel -s 9 -c 3 -f -o gpurun_out/${TAG}_prof_attn_fwd $CMD >> gpurun_out/${TAG}_prof_attn.log 2>&1; echo "attn_fwd rc=$?"
