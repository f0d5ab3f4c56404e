#!/bin/bash
# Quick GPU visit: parity tests + one graph bench.  Usage: gpurun -- 'bash tools/gpu_quick.sh tag [extra command]'
TAG=${1:-q}
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${TAG}_pytest.log
timeout 240 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/${TAG}_bench.err
