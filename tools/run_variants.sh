#!/bin/bash
# tools/run_variants.sh TAG name... -- on the GPU box: bench config 5 with each tuning build tools/libca_<name>.so
tag=$1; shift
out=gpurun_out/${tag}_variants.txt
: > $out
for v in "$@"; do
  CLUSTASSESS_B200_LIB=$PWD/tools/libca_$v.so timeout 300 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline 2> gpurun_out/var_err_$v.txt | grep '^{' | python tools/benchline.py "$v" >> $out
done
cat $out
