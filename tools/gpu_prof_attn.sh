#!/bin/bash
# ncu --set full of the attention kernels , one launch each.  Usage: gpurun -- 'bash tools/gpu_prof_attn.sh tag'
TAG=${1:-r05}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu --graph 0 --sustain 0"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_kernel -s 9 -c 3 -f -o gpurun_out/${TAG}_prof_attn_fwd $CMD > gpurun_out/${TAG}_prof_attn.log 2>&1; echo "fwd rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_kernel -s 9 -c 3 -f -o gpurun_out/${TAG}_prof_attn_bwd $CMD >> gpurun_out/${TAG}_prof_attn.log 2>&1; echo "bwd rc=$?"
ls -la gpurun_out/${TAG}_prof_attn*
