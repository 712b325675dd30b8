#!/bin/bash
# round 2, GPU session ZA (1 GPU, final tree): the whole GPU suite as the driver runs it, after the facet-tagging refactor
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( time timeout 900 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider ) > gpurun_out/r2za_gpu_suite.log 2>&1
echo "suite exit $?" >> gpurun_out/r2za_gpu_suite.log
tail -6 gpurun_out/r2za_gpu_suite.log | cut -c1-250
