#!/bin/bash
# GPU parity suite only.  Usage: gpurun -- bash tools/gpu_pytest.sh tag [pytest args]
TAG=${1:-t}; shift
mkdir -p gpurun_out
python -m pytest tests -m gpu -q "$@" > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest_gpu.log
tail -40 gpurun_out/${TAG}_pytest_gpu.log
