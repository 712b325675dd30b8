#!/bin/bash
# round 2, GPU session Y (1 GPU, final in-tree build: four cases per lane by default in the sweep kernels): the sweep bench lines
# (1024 and 4096 cases), the whole GPU suite as the driver runs it, smoke(), the default bench line
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( time timeout 300 python bench.py --workload sweep --steps 100 --warmup 10 ) > gpurun_out/r2y_bench_sweep.json 2> gpurun_out/r2y_bench_sweep.err
( time timeout 300 python bench.py --workload sweep --cases 4096 --steps 50 --warmup 10 ) > gpurun_out/r2y_bench_sweep_4096.json 2> gpurun_out/r2y_bench_sweep_4096.err
grep -o '"value": [0-9.e+]*\|"ms_per_step": [0-9.]*' gpurun_out/r2y_bench_sweep.json gpurun_out/r2y_bench_sweep_4096.json | tr '\n' ' '; echo
( time timeout 900 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider ) > gpurun_out/r2y_gpu_suite.log 2>&1
echo "suite exit $?" >> gpurun_out/r2y_gpu_suite.log
tail -6 gpurun_out/r2y_gpu_suite.log | cut -c1-250
( time python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r2y_smoke.log 2>&1
tail -4 gpurun_out/r2y_smoke.log | head -1
( time timeout 400 python bench.py --steps 20 --warmup 3 ) > gpurun_out/r2y_bench_default.json 2> gpurun_out/r2y_bench_default.err
grep -o '"ms_per_step": [0-9.]*\|"frac": [0-9.]*' gpurun_out/r2y_bench_default.json | head -3
du -sh gpurun_out
