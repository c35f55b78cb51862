#!/bin/bash
# ncu --set full capture of the morphology kernels (one launch of each) at S=64
mkdir -p gpurun_out
tag=${1:-ncu}
export LT_MORPH_BANDS=${2:-2,5}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_morph -s 8 -c 4 -o gpurun_out/${tag} -f python tools/morph_bench.py --one > gpurun_out/${tag}.log 2>&1
echo "ncu rc=$?"
tail -5 gpurun_out/${tag}.log
ls -la gpurun_out/${tag}.ncu-rep
