#!/bin/bash
# One gpurun call that produces everything profiles/<tag>_* is made from:
#   tools/profile_round.sh <tag> [tests]      (tests: "all" | "none" | pytest selection)
# 1. GPU parity tests  2. the driver's default bench line  3. the ncu launch list of the same command
# 4. one `ncu --set full` capture of a whole warm step (S=64, rendered frames, every kernel once)
set -x
tag=${1:-r02}
tests=${2:-all}
mkdir -p gpurun_out
if [ "$tests" = "all" ]; then
  timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1
  echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
elif [ "$tests" != "none" ]; then
  timeout 1500 python -m pytest $tests -q > gpurun_out/${tag}_pytest.log 2>&1
  echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
fi
tail -8 gpurun_out/${tag}_pytest.log
timeout 900 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
tail -c 400 gpurun_out/${tag}_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_launches_run.log 2>&1
export LT_BENCH_SYNTH=1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_ --launch-skip ${LT_NCU_SKIP:-60} -c ${LT_NCU_COUNT:-22} -f \
  -o gpurun_out/${tag}_step python tools/morph_bench.py --one > gpurun_out/${tag}_ncu_step.log 2>&1
tail -3 gpurun_out/${tag}_ncu_step.log
ls -la gpurun_out/${tag}_*
