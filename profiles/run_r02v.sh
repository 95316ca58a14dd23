# round 2, session v: 256 x 384 tiles in the wgrad pair kernel: parity + A/B kbench + bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_block_gpu.py -m gpu -x -q > gpurun_out/r02v_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 5 gpurun_out/r02v_pytest.log
for w in 0 1; do
CNX_WGRAD_BN384=$w timeout 300 python profiles/kbench.py --only gemm --stages 1,2,3 --iters 5 2>&1 | grep wgrad > gpurun_out/r02v_kbench_wgrad_bn384_$w.jsonl; echo "bn384=$w"; cat gpurun_out/r02v_kbench_wgrad_bn384_$w.jsonl
done
timeout 300 python bench.py --no-cpu-baseline --kernels-out gpurun_out/r02v_kernels.json > gpurun_out/r02v_bench.json 2> gpurun_out/r02v_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r02v_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['roofline']['cnx_kernels_ms_per_step'], d['variants'], d['gpu_launches'])
k=json.load(open('gpurun_out/r02v_kernels.json'))
for f in k['families_all'][:12]: print(f['family'], f['ms'], f['bound'], f['frac'])
for r in k['kernels']:
    if r['family']=='wgrad': print(r['kernel'], r['ms_per_step'], r['tensor_frac'], r['hbm_frac'])"
