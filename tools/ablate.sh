#!/bin/bash
# Run the CA_ABLATE sweep on a GPU box (timings only; results are wrong by construction when a stage is skipped).
export CLUSTASSESS_B200_LIB=$PWD/tools/libclustassess_b200_ablate.so
for d in 0 1 2 3 4 8 7 15 256 512 768; do
  CA_ABLATE=$d python bench.py --config C5 --m ${1:-64} --steps 2 --warmup 1 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('CA_ABLATE=$d rowblock_ms %.2f' % d['stage_ms_per_step']['rowblock_ms'])"
done
