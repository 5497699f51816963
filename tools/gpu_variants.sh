#!/bin/bash
# A/B timing of library variants (development aid): args = variant names under minisat-rust_b200/variants, or "base"
for v in "$@"; do
  echo "== $v"
  if [ "$v" = base ]; then unset B200SAT_LIB; else export B200SAT_LIB=/root/repo/minisat-rust_b200/variants/$v.so; fi
  timeout 100 python tools/gpu_one.py 10000 100 0 0 2>&1 | tail -1 | sed "s/.*last_kernel_ms/ n100 full ms/"
  timeout 100 python tools/gpu_one.py 40000 100 0 0 2>&1 | tail -1 | sed "s/.*last_kernel_ms/ n100 x40k ms/"
done
