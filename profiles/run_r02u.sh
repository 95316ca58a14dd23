# round 2, session u: grouped weight-prep + dz hand-off folded into the dwconv backward-data kernel: parity + bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_dwconv_ln_gpu.py tests/test_block_gpu.py tests/test_patchify_gpu.py tests/test_engine_gpu.py tests/test_ddp_nccl_gpu.py -m gpu -x -q > gpurun_out/r02u_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 6 gpurun_out/r02u_pytest.log
timeout 300 python bench.py --no-cpu-baseline --kernels-out gpurun_out/r02u_kernels.json > gpurun_out/r02u_bench.json 2> gpurun_out/r02u_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r02u_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['roofline']['cnx_kernels_ms_per_step'], d['variants'], d['gpu_launches'])
k=json.load(open('gpurun_out/r02u_kernels.json'))
for f in k['families_all'][:30]: print(f['family'], f['ms'], f['bound'], f['frac'])"
CNX_DZ_HANDOFF=0 timeout 300 python bench.py --no-cpu-baseline --kernels-out gpurun_out/r02u_kernels_nohandoff.json > gpurun_out/r02u_bench_nohandoff.json 2> gpurun_out/r02u_bench_nohandoff.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r02u_bench_nohandoff.json').read().strip().splitlines()[-1])
print('no handoff', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['roofline']['cnx_kernels_ms_per_step'])"
