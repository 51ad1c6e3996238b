#!/bin/sh
# dev helper: build a variant of one kernel file with extra -D flags into build_variants/<name>.so
#   scripts/build_variant.sh <name> <file.cu> <flags...>   (then CYL_LIB=build_variants/<name>.so python scripts/time_hbm.py)
set -e
name=$1; src=$2; shift 2
mkdir -p build_variants
extra=""
[ "$src" = "k_replay.cu" ] && extra="-fmad=false"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr $extra "$@" \
  -c cylinder_b200/csrc/$src -o build_variants/$name.o
objs=$(ls cylinder_b200/build/*.o | grep -v "/${src%.cu}.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build_variants/$name.so build_variants/$name.o $objs -Xcompiler -fPIC -cudart static -ldl -lpthread
rm build_variants/$name.o
