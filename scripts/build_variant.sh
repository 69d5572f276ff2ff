#!/bin/bash
# Builds a variant of the library with extra defines into _variants/libfse_<name>.so (profiling experiments; FSE_B200_LIB selects it).
# usage: scripts/build_variant.sh <name> <nvcc flags...>
set -e
name=$1; shift
cd "$(dirname "$0")/../falling_sand_engine_b200/csrc"
mkdir -p ../../_variants
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC,-fvisibility=hidden "$@" -c fse_tick.cu -o ../../_variants/${name}_tick.o
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC,-fvisibility=hidden "$@" -c fse_capi.cu -o ../../_variants/${name}_capi.o
objs=$(ls _build/*.o | grep -v "fse_tick.o\|fse_capi.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../_variants/libfse_${name}.so ../../_variants/${name}_tick.o ../../_variants/${name}_capi.o $objs -lcudart
echo built _variants/libfse_${name}.so
