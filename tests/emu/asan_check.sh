#!/bin/bash
# TEST-ONLY development aid: run the device solver source under AddressSanitizer inside the SIMT emulator
# (compute-sanitizer is closed on this GPU pool).  Catches out-of-slab / out-of-shared-memory accesses of the
# core, SimpSolver, retry-ladder and portfolio paths on the CPU.   usage: tests/emu/asan_check.sh
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
g++ -std=c++17 -O1 -g -fsanitize=address -fno-omit-frame-pointer -ffp-contract=off -fPIC -DB200SAT_HAVE_SIMP=1 \
    -I"$HERE" -shared -o /tmp/libemu_asan.so "$HERE/emu_harness.cpp"
LD_PRELOAD="$(g++ -print-file-name=libasan.so)" ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0 \
    python "$HERE/asan_run.py"
