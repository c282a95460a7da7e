#!/bin/bash
# Developer tool: builds an experiment variant of libcrb200.so with extra -D flags into build/variants/<name>/.
# usage: tools/build_variant.sh <name> "<extra nvcc flags>"     then   CRB200_LIBRARY=build/variants/<name>/libcrb200.so python bench.py
set -e
cd "$(dirname "$0")/.."
name=$1; extra=$2
out=build/variants/$name
mkdir -p $out
ARCH="-gencode arch=compute_100a,code=sm_100a"
FLAGS="$ARCH -O3 -std=c++17 -lineinfo --extended-lambda -Xcompiler -fPIC -Xcudafe --diag_suppress=177 $extra"
for f in Context Binning BuiltinPipes Resolve; do
  nvcc $FLAGS -c -o $out/$f.o cudaraster-linux_b200/csrc/$f.cu &
done
wait
nvcc $ARCH -shared -Xlinker -Bsymbolic -o $out/libcrb200.so $out/Context.o $out/Binning.o $out/BuiltinPipes.o $out/Resolve.o -ldl
rm -f $out/*.o
echo built $out/libcrb200.so
