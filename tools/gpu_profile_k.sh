#!/bin/bash
# ncu --set full of one kernel by regex.  Usage: bash tools/gpu_profile_k.sh tag name regex skip
TAG=$1; NAME=$2; RX=$3; SKIP=${4:-4}
timeout 300 ncu --set full --clock-control none --import-source on -k regex:$RX -s $SKIP -c 1 -f -o gpurun_out/${TAG}_prof_${NAME} python bench.py --steps 1 --warmup 3 --no-cpu --graph 0 > gpurun_out/${TAG}_prof_${NAME}.log 2>&1; echo "$NAME rc=$?"
