# round 2, session z5: fused split-operand MLP forward (C = 96): parity (own timeout: a barrier bug would hang), kbench, bench
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_block_gpu.py -m gpu -x -q -k "fused_x3" > gpurun_out/r02z5_pytest_a.log 2>&1; echo "pytest(fused x3) rc=$?"; tail -n 15 gpurun_out/r02z5_pytest_a.log
timeout 900 python -m pytest tests/test_block_gpu.py tests/test_engine_gpu.py tests/test_parity_round2_gpu.py -m gpu -x -q > gpurun_out/r02z5_pytest_b.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/r02z5_pytest_b.log
timeout 300 python profiles/kbench.py --only gemm --stages 0 --iters 5 2>&1 | grep "x3" > gpurun_out/r02z5_kbench.jsonl; cut -c1-120 gpurun_out/r02z5_kbench.jsonl
timeout 300 python bench.py --no-cpu-baseline --kernels-out gpurun_out/r02z5_kernels.json > gpurun_out/r02z5_bench.json 2> gpurun_out/r02z5_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r02z5_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['roofline']['cnx_kernels_ms_per_step'], d['variants'], d['gpu_launches'])
k=json.load(open('gpurun_out/r02z5_kernels.json'))
for f in k['families_all'][:14]: print(f['family'], f['ms'], f['bound'], f['frac'])"
