#!/bin/bash
# ncu --set full captures of the dominant kernels (one launch each).  Usage: gpurun -- 'bash tools/gpu_profile.sh tag'
TAG=${1:-r01}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu --graph 0"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_conv_wgrad_kernel -s 8 -c 1 -f -o gpurun_out/${TAG}_prof_wgrad1 $CMD > gpurun_out/${TAG}_prof.log 2>&1; echo "wgrad rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_conv_fprop_kernel -s 15 -c 1 -f -o gpurun_out/${TAG}_prof_fprop1 $CMD >> gpurun_out/${TAG}_prof.log 2>&1; echo "fprop rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_pack_stream_kernel -s 24 -c 1 -f -o gpurun_out/${TAG}_prof_pack $CMD >> gpurun_out/${TAG}_prof.log 2>&1; echo "pack rc=$?"
ls -la gpurun_out/*.ncu-rep
