#!/bin/bash
# developer helper: bench matrix for each prebuilt library variant (variants/<name>.so, built here with -D tunables)
# usage: tools/run_variants.sh "G8 P0 S12" v0 v1 ...
cases=$1; shift
cp datacompressionfloat_b200/libmrczip_b200.so /tmp/orig.so
for v in "$@"; do
  cp variants/$v.so datacompressionfloat_b200/libmrczip_b200.so
  echo "== $v"
  bash tools/bench_matrix.sh $cases
done
cp /tmp/orig.so datacompressionfloat_b200/libmrczip_b200.so
