# round 2, session k: GELU' recomputed in the fc2 data-gradient kernel at C <= 192 (A/B), full suite, bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_block_gpu.py -m gpu -x -q > gpurun_out/r02k_pytest_gemm.log 2>&1; echo "pytest gemm rc=$?"; tail -n 6 gpurun_out/r02k_pytest_gemm.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02k_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/r02k_pytest.log
for v in 0 192; do
  CNX_RECOMPUTE_C=$v timeout 300 python bench.py --no-cpu-baseline --no-variants --kernels-out gpurun_out/r02k_kernels_rc$v.json > gpurun_out/r02k_bench_rc$v.json 2> gpurun_out/r02k_bench_rc$v.err; echo "bench rc$v rc=$?"
done
python - <<'PY'
import json
for v in (0,192):
    d=json.loads(open(f'gpurun_out/r02k_bench_rc{v}.json').read().strip().splitlines()[-1])
    print('recompute C<=',v, d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['roofline']['cnx_kernels_ms_per_step'])
    k=json.load(open(f'gpurun_out/r02k_kernels_rc{v}.json'))
    for r in k['kernels']:
        if r['kernel'].startswith(('fc1_gelu K','dgrad_fc2_gelu')): print('   ', r['kernel'], r['calls_per_step'], r['ms_per_step'], r['hbm_frac'], r['tensor_frac'])
PY
