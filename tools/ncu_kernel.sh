#!/bin/bash
# ncu --set full capture of kernels matching a regex at S=64 on rendered frames: tools/ncu_kernel.sh <tag> <regex> [count]
mkdir -p gpurun_out
tag=$1; regex=$2; cnt=${3:-2}
export LT_BENCH_SYNTH=1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$regex -s 4 -c $cnt -o gpurun_out/${tag} -f python tools/morph_bench.py --one > gpurun_out/${tag}.log 2>&1
echo "ncu rc=$?"
tail -3 gpurun_out/${tag}.log
