# round 2, session q: GELU epilogue with one-instruction symmetric clamps: parity + kbench + bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_block_gpu.py -m gpu -x -q > gpurun_out/r02q_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/r02q_pytest.log
timeout 300 python profiles/kbench.py --only gemm --stages 0,1,2,3 --iters 5 2>&1 | grep -E "fc1_bias_gelu|mlp_fused" > gpurun_out/r02q_kbench_fc1.jsonl; cat gpurun_out/r02q_kbench_fc1.jsonl
timeout 300 python bench.py --no-cpu-baseline --kernels-out gpurun_out/r02q_kernels.json > gpurun_out/r02q_bench.json 2> gpurun_out/r02q_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r02q_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['roofline']['cnx_kernels_ms_per_step'], d['variants'])
k=json.load(open('gpurun_out/r02q_kernels.json'))
for f in k['families_all'][:12]: print(f['family'], f['ms'], f['bound'], f['frac'])"
