#!/bin/bash
# tools/profile_c5.sh TAG -- on the GPU box: the default bench line, the ncu launch list of the same command and one
# `ncu --set full` capture of the wave kernel's two launches (streaming, resident) on config 5.  Outputs in gpurun_out/.
tag=${1:-r1}
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err || exit 1
tail -c 600 gpurun_out/${tag}_bench_n1.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/${tag}_ncu_launches.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:wave_kernel --launch-skip 2 -c 2 -f -o gpurun_out/${tag}_wave \
  python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/${tag}_ncu_full.log 2>&1
ls -la gpurun_out/${tag}_*
