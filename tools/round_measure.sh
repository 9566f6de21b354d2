#!/bin/bash
# tools/round_measure.sh TAG -- on the GPU box: everything profiles/ is built from (1 GPU): the default bench line, ncu launch
# list + full capture (tools/profile_c5.sh), the reference arm, the ablation of the tuning build, the uniform family and the
# other configs' shapes.
tag=${1:-r1}
bash tools/profile_c5.sh $tag
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_reference_arm.json 2> gpurun_out/${tag}_reference_arm.err
ABL_BITS="0 1 2 4 8 16 32 24 28 31" bash tools/ablate_c5.sh $tag > /dev/null
python bench.py --family uniform --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/${tag}_bench_uniform.json 2>/dev/null
python bench.py --config C4 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/${tag}_bench_c4shape.json 2>/dev/null
python bench.py --config C3 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_c3shape.json 2>/dev/null
python tools/run_c3_c4.py > gpurun_out/${tag}_c3_c4.txt 2>&1
python tools/run_c1_c2_dedup.py > gpurun_out/${tag}_c1_c2_dedup.txt 2>&1
nproc > gpurun_out/${tag}_nproc.txt
ls gpurun_out | grep $tag
