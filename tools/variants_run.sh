#!/bin/bash
# tools/variants_run.sh TAG NAME... -- on the GPU box: for every tuning build tools/libca_NAME.so a quick parity subset and the C5 bench line
tag=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  export CLUSTASSESS_B200_LIB=$PWD/tools/libca_$v.so
  timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -m gpu -x -q > gpurun_out/${tag}_${v}_pytest.txt 2>&1
  tail -n 2 gpurun_out/${tag}_${v}_pytest.txt
  timeout 600 python bench.py --steps 4 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/${tag}_${v}_bench.json 2> gpurun_out/${tag}_${v}_bench.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/${tag}_${v}_bench.json"))
    print("$v", "ms/step %.2f" % d["ms_per_step"], "engine %.2f" % d["stage_ms_per_step"]["engine_ms"], "ecc_mean", repr(d["ecc_mean"]), d["stage_ms_per_step"])
except Exception as e:
    print("$v", "FAILED", e)
PY
done
unset CLUSTASSESS_B200_LIB
