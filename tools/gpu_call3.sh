#!/bin/bash
cd /root/repo; mkdir -p gpurun_out
timeout 100 python tools/gpu_k2.py > gpurun_out/k2_plain.log 2>&1; tail -1 gpurun_out/k2_plain.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:replay_smem -s 1 -c 1 -f -o gpurun_out/r02_prof_k2 python tools/gpu_k2.py > gpurun_out/ncu_k2.log 2>&1; tail -2 gpurun_out/ncu_k2.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:search_smem -s 1 -c 1 -f -o gpurun_out/r02_prof_final python tools/gpu_one.py 10000 100 0 0 > gpurun_out/ncu_final.log 2>&1; tail -2 gpurun_out/ncu_final.log
