#!/bin/bash
# One gpurun call: selected parity tests, then timing tools.  Outputs under gpurun_out/.
mkdir -p gpurun_out
tag=${1:-s}
tests=${2:-"tests/test_gpu_config_parity.py::test_tophat_band_geometries tests/test_gpu_parity.py::test_filter_masks_bit_exact"}
timeout 900 python -m pytest $tests -x -q > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${tag}_pytest.log
tail -12 gpurun_out/${tag}_pytest.log
shift; shift
if [ $# -gt 0 ]; then
  timeout 900 "$@" > gpurun_out/${tag}_tool.log 2>&1
  echo "tool rc=$?"
  tail -40 gpurun_out/${tag}_tool.log
fi
