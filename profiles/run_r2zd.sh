#!/bin/bash
# round 2, GPU session ZD (1 GPU, final tree): the reference's drivers end to end - main_wave_2D.py (default case array: the P3 x RT3
# analytical cases of main_wave_2D.py:104-114), main_wave_2D_control.py (passivity + casimir), main_eigen.py (P2 x RT2, nx = 5 .. 25)
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( time timeout 400 python main_wave_2D.py ) > gpurun_out/r2zd_main_wave_2D.log 2>&1; echo "exit $?" >> gpurun_out/r2zd_main_wave_2D.log
( time timeout 400 python main_wave_2D_control.py ) > gpurun_out/r2zd_main_wave_2D_control.log 2>&1; echo "exit $?" >> gpurun_out/r2zd_main_wave_2D_control.log
( time timeout 400 python main_eigen.py ) > gpurun_out/r2zd_main_eigen.log 2>&1; echo "exit $?" >> gpurun_out/r2zd_main_eigen.log
for f in main_wave_2D main_wave_2D_control main_eigen; do echo "== $f"; tail -12 gpurun_out/r2zd_$f.log | cut -c1-200; done
