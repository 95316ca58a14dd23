# round 2, session y2: epilogue chunk round-robin continued across tiles (BN = 192, 16 warps) + L2 prefetch for the split-output fc1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_block_gpu.py -m gpu -x -q > gpurun_out/r02y2_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/r02y2_pytest.log
for w in 0 1; do
CNX_GEMM_PF=$w timeout 300 python profiles/kbench.py --only gemm --stages 0,1 --iters 5 2>&1 | grep "x3\|dgrad_fc2\|fc1_bias" > gpurun_out/r02y2_kbench_pf$w.jsonl; echo "pf=$w"; cut -c1-110 gpurun_out/r02y2_kbench_pf$w.jsonl
done
timeout 300 python bench.py --no-cpu-baseline --kernels-out gpurun_out/r02y2_kernels.json > gpurun_out/r02y2_bench.json 2> gpurun_out/r02y2_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r02y2_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['roofline']['cnx_kernels_ms_per_step'], d['variants'], d['gpu_launches'])
k=json.load(open('gpurun_out/r02y2_kernels.json'))
for f in k['families_all'][:12]: print(f['family'], f['ms'], f['bound'], f['frac'])"
