#!/bin/bash
# Development aid: builds libpfx_b200.so variants that differ in the -D flags of ONE source file, into gpurun_variants/
# usage: tools/build_variants.sh file.cu name1 "-DA=1 -DB=2" name2 "..." ...
set -e
cd "$(dirname "$0")/../pcl_feature_extraction_b200/csrc"
src=$1; shift
base=${src%.cu}
mkdir -p ../../gpurun_variants
others=$(ls ../build/*.o | grep -v "/$base.o")
extra=""
case $base in narf|strict) extra="-fmad=false";; esac
while [ $# -gt 0 ]; do
  name=$1; flags=$2; shift 2
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC $extra $flags -Xptxas -v -c $src -o /tmp/var_$name.o 2> /tmp/var_$name.log
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../gpurun_variants/lib_$name.so $others /tmp/var_$name.o -ldl
  echo "built $name ($flags)"
done
