# round 2, session z2: staged GEMM kernel removed (slab kernel everywhere), 8 epilogue warps for the tensor-bound split-output fc1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_block_gpu.py tests/test_engine_gpu.py -m gpu -x -q > gpurun_out/r02z2_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/r02z2_pytest.log
timeout 300 python bench.py --no-cpu-baseline --kernels-out gpurun_out/r02z2_kernels.json > gpurun_out/r02z2_bench.json 2> gpurun_out/r02z2_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r02z2_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['roofline']['cnx_kernels_ms_per_step'], d['variants'], d['gpu_launches'])
k=json.load(open('gpurun_out/r02z2_kernels.json'))
for f in k['families_all'][:14]: print(f['family'], f['ms'], f['bound'], f['frac'])"
