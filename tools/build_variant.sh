#!/bin/bash
# tools/build_variant.sh NAME FILE.cu [-Dflags...]: link a copy of libds_b200.so whose FILE.cu object is
# compiled with extra flags, into build/variants/NAME.so (select it at run time with DS_B200_LIB=...).
set -e
cd "$(dirname "$0")/.."
name=$1; src=$2; shift 2
OBJDIR=$(python -c "from distantspeech_b200 import _build; print(_build.OBJDIR)")
mkdir -p build/variants
base=$(basename "$src" .cu)
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr \
  -Xptxas -v "$@" -c "distantspeech_b200/csrc/$base.cu" -o "build/variants/$name.$base.o" 2>&1 | grep -A2 "fused\|$base" | grep "spill\|Used" | tail -4
objs=$(ls $OBJDIR/*.o | grep -v "/$base.o")
nvcc -shared -o "build/variants/$name.so" $objs "build/variants/$name.$base.o" -gencode arch=compute_100a,code=sm_100a -lcudart
echo "built build/variants/$name.so"
