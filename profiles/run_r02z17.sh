# round 2, session z17: fp32 training with GELU, split, GELU'(h) fused into the fc1 epilogue and the GELU' multiply + split into the fc2 data-gradient epilogue: parity + fp32 bench on / off
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_block_gpu.py tests/test_engine_gpu.py tests/test_gemm_gpu.py tests/test_parity_round2_gpu.py tests/test_ddp_nccl_gpu.py -m gpu -x -q > gpurun_out/r02z17_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/r02z17_pytest.log
for v in 1 0; do
CNX_X3_TRAIN_FUSED=$v timeout 300 python bench.py --no-amp --no-cpu-baseline --no-variants --kernels-out gpurun_out/r02z17_kernels_fp32_fused$v.json > gpurun_out/r02z17_bench_fp32_fused$v.json 2> gpurun_out/r02z17_bench.err; python -c "
import json
d=json.loads(open('gpurun_out/r02z17_bench_fp32_fused$v.json').read().strip().splitlines()[-1])
print('fp32 fused=$v', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['roofline']['cnx_kernels_ms_per_step'])
k=json.load(open('gpurun_out/r02z17_kernels_fp32_fused$v.json'))
for f in k['families_all'][:9]: print('   ', f['family'], f['ms'], f['bound'], f['frac'])"
done
