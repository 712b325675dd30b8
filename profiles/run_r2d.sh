#!/bin/bash
# round 2, GPU session D (2 GPUs): where does the per-iteration cost of the inter-GPU exchange go?  Small slabs (3.1 M dofs per GPU:
# a sweep takes ~10-20 us, so the exchange dominates), A/B knobs PHW_XCH_DEBUG (bit0: no system fence after the mid-sweep pushes,
# bit1: no consumer-side fence, bit3: no producer-side fence, bit2/bit4: gpu-scope instead of system-scope fences, bit5: round-1 protocol)
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
ARGS="--nx 1768 --ny 442 --no-strong --steps 100 --warmup 10 --no-cpu-baseline"
summ() { grep '^{' $1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; c=d['config']
print('$2', 'gpus', d['n_gpus'], 'ms/step %.4f' % d['ms_per_step'], 'us/iter p %.2f q %.2f' % (1e3*r['ms_per_iter_p'], 1e3*r['ms_per_iter_q']), 'its %.2f %.2f' % (c['iters_per_solve_p'], c['iters_per_solve_q']), 'res %.2e' % c['max_rel_residual'], 'err %.2e' % c['exact_error']['p_nodal_rel_l2'])
" >> gpurun_out/r2d_exchange_ab.txt; }
( CUDA_VISIBLE_DEVICES=0 timeout 120 python bench.py --gpus 1 $ARGS ) > gpurun_out/r2d_1gpu.log 2>&1; summ gpurun_out/r2d_1gpu.log "1gpu            "
for dbg in 0 32 1 2 8 11 20; do
  ( PHW_XCH_DEBUG=$dbg timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 $ARGS ) > gpurun_out/r2d_dbg$dbg.log 2>&1
  summ gpurun_out/r2d_dbg$dbg.log "PHW_XCH_DEBUG=$dbg"
done
( PHW_PIPE_MULTI=0 timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 $ARGS --pipe none ) > gpurun_out/r2d_lean.log 2>&1
summ gpurun_out/r2d_lean.log "lean kernels    "
( CUDA_VISIBLE_DEVICES=0 timeout 120 python bench.py --gpus 1 $ARGS --pipe none ) > gpurun_out/r2d_1gpu_lean.log 2>&1; summ gpurun_out/r2d_1gpu_lean.log "1gpu lean       "
cat gpurun_out/r2d_exchange_ab.txt
