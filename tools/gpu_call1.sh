#!/bin/bash
# A/B of library variants + two ncu source-level captures (saturated n=100 batch; one warp per SM at n=250)
cd /root/repo; mkdir -p gpurun_out
bash tools/gpu_ab.sh base opq opq24 > gpurun_out/r02_ab_d.log 2>&1
cat gpurun_out/r02_ab_d.log
export B200SAT_LIB=/root/repo/minisat-rust_b200/variants/opq.so
timeout 120 python tools/gpu_one.py 148 250 0 0 3000 > gpurun_out/lone_plain.log 2>&1; tail -1 gpurun_out/lone_plain.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:search_smem -s 1 -c 1 -f -o gpurun_out/r02_prof_lone python tools/gpu_one.py 148 250 0 0 3000 > gpurun_out/ncu_lone.log 2>&1; tail -2 gpurun_out/ncu_lone.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:search_smem -s 1 -c 1 -f -o gpurun_out/r02_prof_opq python tools/gpu_one.py 10000 100 0 0 > gpurun_out/ncu_opq.log 2>&1; tail -2 gpurun_out/ncu_opq.log
ls -la gpurun_out | tail -8
