#!/bin/bash
# tools/ablate_c5.sh TAG -- on the GPU box: the wave kernel with parts switched off (tuning build tools/libca_abl.so,
# built with tools/build_variant.sh abl -DCA_ABLATE), config 5, kernel ms per step.
# bits: 1 no ecc flush, 2 no table fill, 4 no label loads (all labels 0), 8 no histogram, 16 no gather
tag=${1:-r1}
out=gpurun_out/${tag}_ablate.txt
echo "# CA_ABLATE bits: 1 no ecc flush, 2 no table fill, 4 no label loads, 8 no histogram, 16 no gather" > $out
for bits in ${ABL_BITS:-0 1 2 4 8 16 24 28 31}; do
  CA_ABLATE=$bits CLUSTASSESS_B200_LIB=$PWD/tools/libca_abl.so timeout 300 python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline 2> gpurun_out/abl_err_$bits.txt | tee gpurun_out/abl_out_$bits.txt | grep '^{' | python tools/benchline.py "CA_ABLATE=$bits" >> $out; grep '^.kind' gpurun_out/abl_out_$bits.txt | sort | uniq | tail -8 >> $out
done
cat $out
