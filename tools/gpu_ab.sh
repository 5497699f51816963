#!/bin/bash
# A/B timing aid: config-2 shape at two batch sizes + a bounded n=250 slice; prints launch info and props/s
for v in "$@"; do
  echo "== $v"
  if [ "$v" = base ]; then unset B200SAT_LIB; else export B200SAT_LIB=/root/repo/minisat-rust_b200/variants/$v.so; fi
  timeout 200 python tools/gpu_one.py 10000 100 0 0 2>&1 | tail -1
  timeout 200 python tools/gpu_one.py 40000 100 0 0 2>&1 | tail -1
  timeout 300 python tools/gpu_one.py 4736 250 0 0 20000 2>&1 | tail -1
done
