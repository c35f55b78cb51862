#!/bin/bash
mkdir -p gpurun_out
tag=${1:-full}
./tools/_build/pipe_probe2 > gpurun_out/${tag}_pipe_probe2.jsonl 2>&1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -25 gpurun_out/${tag}_pytest.log
cat gpurun_out/${tag}_pipe_probe2.jsonl
