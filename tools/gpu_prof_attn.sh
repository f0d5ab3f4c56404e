#!/bin/bash
# ncu --set full of the forward attention kernels, one launch each.  Usage: gpurun -- 'bash tools/gpu_prof_attn.sh tag [kernel regex]'
TAG=${1:-r06}
RE=${2:-attn_fwd}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu --graph 0 --sustain 0"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:$RE -s 9 -c 3 -f -o gpurun_out/${TAG}_prof_attn $CMD > gpurun_out/${TAG}_prof_attn.log 2>&1; echo "rc=$?"
ls -la gpurun_out/${TAG}_prof_attn*
