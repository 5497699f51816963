#!/bin/bash
# round-end validation on one B200: GPU test suite, smoke, both bench arms, ncu launch list of the bench
cd /root/repo; mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -s > gpurun_out/r02_pytest_final.log 2>&1; tail -3 gpurun_out/r02_pytest_final.log
timeout 90 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; tail -1 gpurun_out/r02_smoke.log | cut -c1-200
timeout 300 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/bench.err; tail -c 400 gpurun_out/bench.err; cut -c1-400 gpurun_out/r02_bench_final.json
timeout 200 python bench.py --pipeline 4 --no-cpu-baseline --no-bcp-replay > gpurun_out/r02_bench_pipeline4.json 2>/dev/null; cut -c1-200 gpurun_out/r02_bench_pipeline4.json
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_reference.json 2>/dev/null; cut -c1-200 gpurun_out/r02_bench_reference.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; tail -2 gpurun_out/ncu_bench.log | cut -c1-300
