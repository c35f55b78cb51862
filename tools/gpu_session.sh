#!/bin/bash
# One gpurun call of the round: parity tests, the pipe probe, a short bench.  Outputs under gpurun_out/.
mkdir -p gpurun_out
tag=${1:-s1}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.txt 2>&1
./tools/_build/pipe_probe > gpurun_out/${tag}_pipe_probe.jsonl 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -15 gpurun_out/${tag}_pytest.log
python bench.py --steps 100 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench rc=$?"
cat gpurun_out/${tag}_pipe_probe.jsonl
