#!/bin/bash
# tools/round_measure.sh TAG -- on the GPU box: everything profiles/ is built from (1 GPU): the default bench line (with the parity
# block and the cpu baseline), ncu launch list + full capture (tools/profile_c5.sh), the reference arm, the uniform family, the other
# configs' shapes (C4 as element_sim_matrix, 8 B per unit), the small-shape latencies and the dedup pre-pass.
tag=${1:-r2}
bash tools/profile_c5.sh $tag
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_reference_arm.json 2> gpurun_out/${tag}_reference_arm.err
timeout 300 python bench.py --family uniform --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/${tag}_bench_uniform.json 2>/dev/null
timeout 300 python bench.py --config C4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_c4shape.json 2>/dev/null
timeout 300 python bench.py --config C4 --mode sim_matrix --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_c4_sim_matrix.json 2>/dev/null
timeout 300 python bench.py --config C3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_c3shape.json 2>/dev/null
timeout 300 python tools/run_c3_c4.py > gpurun_out/${tag}_c3_c4.txt 2>&1
timeout 600 python tools/run_c1_c2_dedup.py > gpurun_out/${tag}_c1_c2_dedup.txt 2>&1
nproc > gpurun_out/${tag}_nproc.txt
ls gpurun_out | grep $tag
