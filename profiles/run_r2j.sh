#!/bin/bash
# round 2, GPU session J (2 GPUs): partitioned-run parity (lean + pipelined) with the new layout (halo-only patches last, not processed),
# the bench under torchrun (weak + strong), and on GPU 0 the sweep with pipelined solves on the grouped mesh
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( time PHW_TEST_PIPE_MULTI=1 timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -p no:cacheprovider -rs ) > gpurun_out/r2j_multi_parity.log 2>&1
echo "multi parity exit $?" >> gpurun_out/r2j_multi_parity.log
tail -8 gpurun_out/r2j_multi_parity.log
( time timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline ) > gpurun_out/r2j_bench_2gpu.log 2>&1
echo "bench exit $?" >> gpurun_out/r2j_bench_2gpu.log
tail -c 2500 gpurun_out/r2j_bench_2gpu.log
for v in "--patch-cells 1024 --sweep-pipe --pipe pq" "--patch-cells 1024 --sweep-pipe --pipe q" "--patch-cells 1024 --pipe none" "--patch-cells 2704 --pipe none"; do
  tag=$(echo "$v" | tr -d ' -'); 
  ( time PHW_NO_LOOP=1 timeout 300 python bench.py --workload sweep --steps 200 --warmup 10 $v ) > gpurun_out/r2j_sweep_$tag.log 2>&1
  echo "sweep [$v]: $(grep -o '"value": [0-9.e+]*' gpurun_out/r2j_sweep_$tag.log | head -1) $(grep -o '"ms_per_step": [0-9.e+]*' gpurun_out/r2j_sweep_$tag.log | head -1) $(grep -o '"iters_per_solve_p": [0-9.]*, "iters_per_solve_q": [0-9.]*' gpurun_out/r2j_sweep_$tag.log) $(grep -o '"kernel_paths": [0-9]*' gpurun_out/r2j_sweep_$tag.log) $(grep -c Error gpurun_out/r2j_sweep_$tag.log) errors"
done
