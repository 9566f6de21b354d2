#!/bin/bash
# tools/build_variant.sh NAME [nvcc flags...] -- builds tools/libca_NAME.so with extra compile flags (tuning experiments);
# run it with CLUSTASSESS_B200_LIB=$PWD/tools/libca_NAME.so python bench.py ...
set -e
name=$1; shift
cd "$(dirname "$0")/.."
out=tools/libca_$name.so
objs=""
mkdir -p /tmp/ca_variant_$name
for f in clustassess_b200/csrc/*.cu clustassess_b200/csrc/*.cpp; do
  o=/tmp/ca_variant_$name/$(basename $f).o
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC,-O3 "$@" -c $f -o $o &
  objs="$objs $o"
done
wait
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $out $objs -lcudart_static -lpthread -ldl -lrt
echo built $out
