#!/bin/bash
# round 2, GPU session ZC (1 GPU, final tree): the reference's own regression harness on this path - run_test_cases.py test_cases (the
# eleven quick cases of run_test_cases.py:23-35 incl. the higher-order analytical ones) compared by test_case_check at the reference's
# thresholds against the CPU oracle on the same meshes
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( time timeout 600 python run_test_cases.py test_cases ) > gpurun_out/r2zc_run_test_cases.log 2>&1
echo "exit $?" >> gpurun_out/r2zc_run_test_cases.log
tail -40 gpurun_out/r2zc_run_test_cases.log | cut -c1-200
