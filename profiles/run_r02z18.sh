# round 2, session z18: wgrad_x3 as one launch with a three-pass K loop: parity + fp32 bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_block_gpu.py tests/test_engine_gpu.py -m gpu -x -q > gpurun_out/r02z18_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 6 gpurun_out/r02z18_pytest.log | cut -c1-200
for v in 1 0; do
CNX_WGRAD_X3_PASSES=$v timeout 300 python bench.py --no-amp --no-cpu-baseline --no-variants --kernels-out gpurun_out/r02z18_kernels_fp32_$v.json > gpurun_out/r02z18_bench_fp32_$v.json 2> gpurun_out/r02z18_bench.err; python -c "
import json
d=json.loads(open('gpurun_out/r02z18_bench_fp32_$v.json').read().strip().splitlines()[-1])
print('fp32 one_loop=$v', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['roofline']['cnx_kernels_ms_per_step'])
k=json.load(open('gpurun_out/r02z18_kernels_fp32_$v.json'))
for f in k['families_all'][:5]: print('   ', f['family'], f['ms'], f['bound'], f['frac'])"
done
