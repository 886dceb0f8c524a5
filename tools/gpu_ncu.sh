#!/bin/bash
# usage: tools/gpu_ncu.sh <tag> [bench args...]   -- one ncu --set full capture of the rx kernel
tag=$1; shift
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_rx -s 1 -c 1 -f -o gpurun_out/$tag \
   python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-configs "$@" > gpurun_out/$tag.log 2>&1
tail -2 gpurun_out/$tag.log | cut -c1-300
