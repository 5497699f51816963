#!/bin/bash
# build a library variant for A/B timing: tools/build_variant.sh NAME "extra nvcc flags"   -> minisat-rust_b200/variants/NAME.so
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
src=$root/minisat-rust_b200/csrc
out=$root/minisat-rust_b200/variants
tmp=$(mktemp -d)
mkdir -p $out
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --fmad=false -Xcompiler -fPIC -DB200SAT_HAVE_SIMP=1 -I$src $*"
# only the hot search kernel is rebuilt with the extra flags; everything else is taken from the base build
nvcc $FLAGS -Xptxas -v -c -o $tmp/kern_search_c0.o $src/kern_search_c0.cu 2> $out/$name.build.log
objs=""
for o in b200sat preprocess_kernels preprocess_coop dimacs_io kern_search_c1 kern_search_pf kern_search_rescue kern_replay_c0 kern_replay_c1 kern_ingest_c0s0 kern_ingest_c1s0 kern_ingest_c0s1 kern_ingest_c1s1; do objs="$objs $src/$o.o"; done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/$name.so $tmp/kern_search_c0.o $objs -lcudart_static -lrt -ldl -lpthread -lz
grep -E "search_smemILi128" -A2 $out/$name.build.log | grep -E "registers|spill" | head -2
rm -rf $tmp
