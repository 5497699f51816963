#!/bin/bash
# A/B timing of library variants (development aid)
for v in "$@"; do
  echo "== $v"
  B200SAT_LIB=/root/repo/minisat-rust_b200/variants/$v.so timeout 100 python tools/gpu_one.py 3000 100 1 2 2>&1 | tail -1 | sed 's/.*last_kernel_ms/low-occ ms/'
  B200SAT_LIB=/root/repo/minisat-rust_b200/variants/$v.so timeout 100 python tools/gpu_one.py 10000 100 0 0 2>&1 | tail -1 | sed "s/'slab_bytes.*last_kernel_ms/ full ms/"
done
