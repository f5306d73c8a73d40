#!/bin/bash
# build a variant of libunit_b200.so with extra -D flags for one source: tools/build_variant.sh NAME SRC "-DX=1 -DY=2"
set -e
cd "$(dirname "$0")/.."
python -m unit_b200.build > /dev/null
mkdir -p unit_b200/build/variants
name=$1; src=$2; flags=$3
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr $flags -c unit_b200/csrc/$src.cu -o unit_b200/build/variants/${src}_$name.o -Iinclude
objs=$(ls unit_b200/build/*.o | grep -v "/$src.o")
nvcc -shared -o unit_b200/build/variants/lib_$name.so $objs unit_b200/build/variants/${src}_$name.o -gencode arch=compute_100a,code=sm_100a
echo unit_b200/build/variants/lib_$name.so
