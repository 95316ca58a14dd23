# round 2, session z13: fused x3 MLP with the final epilogue transposed through shared memory (coalesced residual / output rows)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_block_gpu.py tests/test_engine_gpu.py tests/test_parity_round2_gpu.py -m gpu -x -q > gpurun_out/r02z13_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r02z13_pytest.log
timeout 300 python profiles/kbench.py --only gemm --stages 0 --iters 5 2>&1 | grep "x3" > gpurun_out/r02z13_kbench.jsonl; cut -c1-120 gpurun_out/r02z13_kbench.jsonl
for v in 1 0; do
CNX_FUSED_MLP_X3=$v timeout 300 python bench.py --no-cpu-baseline --no-variants > gpurun_out/r02z13_bench_fused$v.json 2> gpurun_out/r02z13_bench.err; python -c "
import json
d=json.loads(open('gpurun_out/r02z13_bench_fused$v.json').read().strip().splitlines()[-1])
print('fused=$v', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['roofline']['cnx_kernels_ms_per_step'], d['clocks']['sm_mhz'])"
done
