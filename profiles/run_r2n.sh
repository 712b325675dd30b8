#!/bin/bash
# round 2, GPU session N (1 GPU): the whole GPU suite exactly as the driver runs it, the higher-order paper cases with the 16-lanes-per-cell
# matrix-free kernels (and assembled CSR rows beside them, plus one mesh that does not fit the L2), the default bench line, the SV and
# from-rest lines of SURVEY 8d, the paper and sweep workloads, ncu --set full of the matrix-free solve and of the whole-loop kernel
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
# the binary built from the sources of this session, frozen (the in-tree sources may be edited while the call waits in the queue)
[ -f porthamiltonian_fem_b200/libphwave_frozen.so ] && export PHW_LIB="$PWD/porthamiltonian_fem_b200/libphwave_frozen.so"
( time timeout 900 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider ) > gpurun_out/r2n_gpu_suite.log 2>&1
echo "suite exit $?" >> gpurun_out/r2n_gpu_suite.log
tail -6 gpurun_out/r2n_gpu_suite.log
( time CPU_STEPS=20 timeout 400 python profiles/ho_cases.py ) > gpurun_out/r2n_ho_mf.jsonl 2> gpurun_out/r2n_ho_mf.err
( time CPU_STEPS=20 PHW_HO_CSR=1 timeout 400 python profiles/ho_cases.py ) > gpurun_out/r2n_ho_csr.jsonl 2> gpurun_out/r2n_ho_csr.err
( time timeout 300 python profiles/ho_large.py ) > gpurun_out/r2n_ho_large.jsonl 2> gpurun_out/r2n_ho_large.err
python - <<'PY'
import json
for f in ('gpurun_out/r2n_ho_mf.jsonl', 'gpurun_out/r2n_ho_csr.jsonl', 'gpurun_out/r2n_ho_large.jsonl'):
    print(f)
    for ln in open(f):
        if ln.startswith('{'):
            d = json.loads(ln)
            print('  %-46s %8d dofs %8.0f steps/s mf=%s its %.1f %.1f' % (d['case'], d['dofs'], d['gpu_steps_per_s'], d.get('matrix_free'), d['iters_p'], d['iters_q']))
PY
( time timeout 400 python bench.py --steps 20 --warmup 3 ) > gpurun_out/r2n_bench_default.json 2> gpurun_out/r2n_bench_default.err
grep -o '"ms_per_step": [0-9.]*\|"frac": [0-9.]*\|"whole_loop_model_frac": [0-9.]*' gpurun_out/r2n_bench_default.json | head -5
( time timeout 300 python bench.py --scheme SV --steps 10 --warmup 3 --no-cpu-baseline ) > gpurun_out/r2n_bench_sv.json 2> gpurun_out/r2n_bench_sv.err
( time timeout 300 python bench.py --from-rest --steps 10 --warmup 3 --no-cpu-baseline ) > gpurun_out/r2n_bench_from_rest.json 2> gpurun_out/r2n_bench_from_rest.err
grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2n_bench_sv.json gpurun_out/r2n_bench_from_rest.json
( time timeout 400 python bench.py --workload paper ) > gpurun_out/r2n_bench_paper.json 2> gpurun_out/r2n_bench_paper.err
( time timeout 300 python bench.py --workload sweep --steps 100 --warmup 10 ) > gpurun_out/r2n_bench_sweep.json 2> gpurun_out/r2n_bench_sweep.err
grep -o '"value": [0-9.e+]*' gpurun_out/r2n_bench_paper.json gpurun_out/r2n_bench_sweep.json | head -4
( time HO_STEPS=20 CPU_STEPS=2 HO_ONLY=0 timeout 300 ncu --set full --clock-control none --import-source on -k 'regex:ho_solve_k|ho_apply_k|ho_quad_k' -c 8 -f -o gpurun_out/r2n_ho_full python profiles/ho_cases.py ) > gpurun_out/r2n_ncu_ho.log 2>&1
tail -2 gpurun_out/r2n_ncu_ho.log
( time PAPER_ONLY=0 PAPER_STEPS=200 CPU_STEPS=0 timeout 300 ncu --set full --clock-control none --import-source on -k 'regex:loop_k' -c 1 -f -o gpurun_out/r2n_loop_full python profiles/paper_cases.py ) > gpurun_out/r2n_ncu_loop.log 2>&1
tail -2 gpurun_out/r2n_ncu_loop.log
ls -la gpurun_out/*.ncu-rep
