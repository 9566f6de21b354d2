#!/bin/bash
# tools/run_env.sh TAG VAR v1 v2 ... -- on the GPU box: bench config 5 once per value of an environment tuning knob
tag=$1; var=$2; shift; shift
out=gpurun_out/${tag}_env.txt
: > $out
for v in "$@"; do
  env $var=$v timeout 300 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline 2> gpurun_out/env_err_$v.txt | grep '^{' | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$var=$v', 'ms/step %.1f'%d['ms_per_step'], {k:round(x,2) for k,x in d['stage_ms_per_step'].items()})" >> $out
done
cat $out
