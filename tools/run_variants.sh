#!/bin/bash
# Development aid (GPU box): runs a command once per library variant built by tools/build_variants.sh
# usage: tools/run_variants.sh "grep-pattern" cmd...     (the default build runs first as "base")
pat=$1; shift
lib=pcl_feature_extraction_b200/lib/libpfx_b200.so
cp $lib /tmp/lib_base.so
echo "== base"; "$@" 2>&1 | grep -E "$pat"
for v in gpurun_variants/lib_*.so; do
  cp $v $lib
  echo "== $v"; "$@" 2>&1 | grep -E "$pat"
done
cp /tmp/lib_base.so $lib
