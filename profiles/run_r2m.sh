#!/bin/bash
# round 2, GPU session M (1 GPU): matrix-free higher-order kernels (csrc/kernels_ho.cuh) - parity + operator tests, eigenvalue tests, the
# sweep tests with the new default, then the reference's higher-order paper cases matrix-free and (PHW_HO_CSR=1) with assembled CSR rows
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
( time timeout 900 python -m pytest tests/test_gpu_high_order.py tests/test_gpu_eigen.py tests/test_gpu_field_output.py -m gpu -q -p no:cacheprovider ) > gpurun_out/r2m_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r2m_tests.log
tail -30 gpurun_out/r2m_tests.log
( time timeout 300 python -m pytest tests/test_gpu_loop.py -m gpu -q -p no:cacheprovider -k sweep ) > gpurun_out/r2m_sweep_tests.log 2>&1
tail -4 gpurun_out/r2m_sweep_tests.log
( time CPU_STEPS=20 timeout 600 python profiles/ho_cases.py ) > gpurun_out/r2m_ho_mf.jsonl 2> gpurun_out/r2m_ho_mf.err
( time CPU_STEPS=20 PHW_HO_CSR=1 timeout 600 python profiles/ho_cases.py ) > gpurun_out/r2m_ho_csr.jsonl 2> gpurun_out/r2m_ho_csr.err
python - <<'PY'
import json
for f in ('gpurun_out/r2m_ho_mf.jsonl', 'gpurun_out/r2m_ho_csr.jsonl'):
    print(f)
    for ln in open(f):
        if ln.startswith('{'):
            d = json.loads(ln)
            print('  %-46s %7d dofs %8.0f steps/s mf=%s launches=%d its %.1f %.1f setup+run %.1f s cpu %5.1f steps/s diff %.1e' % (d['case'], d['dofs'], d['gpu_steps_per_s'], d.get('matrix_free'), d['launches'], d['iters_p'], d['iters_q'], d['gpu_wall_s_incl_setup'], d['cpu_oracle_steps_per_s'], [v for k, v in d.items() if k.startswith('H_array_rel_diff')][0]))
PY
tail -3 gpurun_out/r2m_ho_mf.err
