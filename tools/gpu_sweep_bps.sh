#!/bin/bash
# occupancy sweep: resident one-warp blocks per SM (more L1 per warp vs more warps)
for bps in 0 28 24 20 16; do echo "n100 bps=$bps"; timeout 200 python tools/gpu_one.py 40000 100 0 $bps 2>&1 | tail -1; done
for bps in 0 20 16 12 8; do echo "n250 bps=$bps"; timeout 300 python tools/gpu_one.py 4736 250 0 $bps 20000 2>&1 | tail -1; done
