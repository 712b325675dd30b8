#!/bin/bash
# round 2, GPU session P (2 GPUs): the whole GPU suite on a 2-GPU box, so that the partitioned parity cases run too - new: implicit schemes
# (IE, SM) and the lumped model (SE / SV / SM + IC) on a partitioned mesh; conjugate-gradient mass solves; higher-order default selection
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
[ -f porthamiltonian_fem_b200/libphwave_frozen.so ] && export PHW_LIB="$PWD/porthamiltonian_fem_b200/libphwave_frozen.so"
nvidia-smi -L > gpurun_out/r2p_gpus.txt
( time timeout 1200 python -m pytest tests/ -q -m gpu -p no:cacheprovider -rs ) > gpurun_out/r2p_gpu_suite_2gpus.log 2>&1
echo "suite exit $?" >> gpurun_out/r2p_gpu_suite_2gpus.log
tail -60 gpurun_out/r2p_gpu_suite_2gpus.log | cut -c1-250
