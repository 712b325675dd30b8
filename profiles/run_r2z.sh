#!/bin/bash
# round 2, GPU session Z (1 GPU, final in-tree build): cyclic row tiles (PHW_SWEEP_MAP=1) on top of four cases per lane, 1024 / 4096 cases
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
for map in 1 0; do
  ( time PHW_SWEEP_MAP=$map timeout 300 python bench.py --workload sweep --steps 100 --warmup 10 ) > gpurun_out/r2z_bench_sweep_map$map.json 2> gpurun_out/r2z_bench_sweep_map$map.err
  ( time PHW_SWEEP_MAP=$map timeout 300 python bench.py --workload sweep --cases 4096 --steps 50 --warmup 10 ) > gpurun_out/r2z_bench_sweep_4096_map$map.json 2> gpurun_out/r2z_bench_sweep_4096_map$map.err
  grep -o '"value": [0-9.e+]*' gpurun_out/r2z_bench_sweep_map$map.json gpurun_out/r2z_bench_sweep_4096_map$map.json | tr '\n' ' '; echo " map=$map"
done
