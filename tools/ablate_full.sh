#!/bin/bash
export CLUSTASSESS_B200_LIB=$PWD/tools/libclustassess_b200_ablate.so
for d in 0 8 3 11 15 256; do
  CA_ABLATE=$d python bench.py --config C5 --m ${1:-500} --steps 2 --warmup 1 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('CA_ABLATE=$d rowblock_ms %.2f' % d['stage_ms_per_step']['rowblock_ms'])"
done
