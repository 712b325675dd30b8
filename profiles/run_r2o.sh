#!/bin/bash
# round 2, GPU session O (1 GPU): the whole GPU suite exactly as the driver runs it (new: higher-order implicit schemes, conjugate-gradient
# mass solves on poor meshes, eigenvalue tests), the higher-order paper cases (matrix-free at three CTA sizes, assembled CSR beside them,
# meshes beyond the L2), the bench lines of SURVEY 8d (default, SV, from rest, paper cases, sweep), ncu --set full of the matrix-free
# solve and of the whole-loop kernel (summarised on the box: the .ncu-rep files are too large to travel)
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
# the binary built from the sources of this session, frozen (the in-tree sources may be edited while the call waits in the queue)
[ -f porthamiltonian_fem_b200/libphwave_frozen.so ] && export PHW_LIB="$PWD/porthamiltonian_fem_b200/libphwave_frozen.so"
( time timeout 900 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider ) > gpurun_out/r2o_gpu_suite.log 2>&1
echo "suite exit $?" >> gpurun_out/r2o_gpu_suite.log
tail -6 gpurun_out/r2o_gpu_suite.log
for nt in 256 512 1024; do
  ( time CPU_STEPS=2 PHW_HO_NT=$nt timeout 300 python profiles/ho_cases.py ) > gpurun_out/r2o_ho_mf_nt$nt.jsonl 2> gpurun_out/r2o_ho_mf_nt$nt.err
done
( time CPU_STEPS=20 timeout 400 python profiles/ho_cases.py ) > gpurun_out/r2o_ho_mf.jsonl 2> gpurun_out/r2o_ho_mf.err
( time CPU_STEPS=2 PHW_HO_CSR=1 timeout 400 python profiles/ho_cases.py ) > gpurun_out/r2o_ho_csr.jsonl 2> gpurun_out/r2o_ho_csr.err
( time timeout 400 python profiles/ho_large.py ) > gpurun_out/r2o_ho_large.jsonl 2> gpurun_out/r2o_ho_large.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2o_ho_*.jsonl')):
    print(f)
    for ln in open(f):
        if ln.startswith('{'):
            d = json.loads(ln)
            print('  %-46s %8d dofs %8.0f steps/s mf=%s its %.1f %.1f' % (d['case'], d['dofs'], d['gpu_steps_per_s'], d.get('matrix_free'), d['iters_p'], d['iters_q']))
PY
( time timeout 400 python bench.py --steps 20 --warmup 3 ) > gpurun_out/r2o_bench_default.json 2> gpurun_out/r2o_bench_default.err
grep -o '"ms_per_step": [0-9.]*\|"frac": [0-9.]*\|"whole_loop_model_frac": [0-9.]*' gpurun_out/r2o_bench_default.json | head -5
( time timeout 300 python bench.py --scheme SV --steps 10 --warmup 3 --no-cpu-baseline ) > gpurun_out/r2o_bench_sv.json 2> gpurun_out/r2o_bench_sv.err
( time timeout 300 python bench.py --from-rest --steps 10 --warmup 3 --no-cpu-baseline ) > gpurun_out/r2o_bench_from_rest.json 2> gpurun_out/r2o_bench_from_rest.err
grep -o '"ms_per_step": [0-9.]*' gpurun_out/r2o_bench_sv.json gpurun_out/r2o_bench_from_rest.json
( time timeout 400 python bench.py --workload paper ) > gpurun_out/r2o_bench_paper.json 2> gpurun_out/r2o_bench_paper.err
( time timeout 300 python bench.py --workload sweep --steps 100 --warmup 10 ) > gpurun_out/r2o_bench_sweep.json 2> gpurun_out/r2o_bench_sweep.err
grep -o '"value": [0-9.e+]*' gpurun_out/r2o_bench_paper.json gpurun_out/r2o_bench_sweep.json | head -4
# ncu: summaries made here, reports deleted (two of them were 100 MB in session N and nothing came back)
( time HO_STEPS=20 CPU_STEPS=2 HO_ONLY=0 timeout 300 ncu --set full --clock-control none -k 'regex:ho_solve_k|ho_apply_k|ho_quad_k' -c 6 -f -o /tmp/r2o_ho_full python profiles/ho_cases.py ) > gpurun_out/r2o_ncu_ho.log 2>&1
python profiles/ncu_extract.py /tmp/r2o_ho_full.ncu-rep gpurun_out/r2o_ncu_full_ho_matrix_free.txt > /dev/null 2>> gpurun_out/r2o_ncu_ho.log
ncu -i /tmp/r2o_ho_full.ncu-rep --page details 2>/dev/null | grep -v "^ *$" | head -700 > gpurun_out/r2o_ncu_ho_details.txt
( time PAPER_ONLY=0 PAPER_STEPS=200 CPU_STEPS=0 timeout 300 ncu --set full --clock-control none -k 'regex:loop_k' -c 1 -f -o /tmp/r2o_loop_full python profiles/paper_cases.py ) > gpurun_out/r2o_ncu_loop.log 2>&1
python profiles/ncu_extract.py /tmp/r2o_loop_full.ncu-rep gpurun_out/r2o_ncu_full_whole_loop_kernel.txt > /dev/null 2>> gpurun_out/r2o_ncu_loop.log
ncu -i /tmp/r2o_loop_full.ncu-rep --page details 2>/dev/null | grep -v "^ *$" | head -300 > gpurun_out/r2o_ncu_loop_details.txt
cat gpurun_out/r2o_ncu_full_ho_matrix_free.txt gpurun_out/r2o_ncu_full_whole_loop_kernel.txt | cut -c1-200
du -sh gpurun_out
