# round 2, session r: LayerNorm half of the fp32-activation dwconv with 8-channel lane items (16-byte split stores): parity + kbench + bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dwconv_ln_gpu.py tests/test_block_gpu.py tests/test_engine_gpu.py -m gpu -x -q > gpurun_out/r02r_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/r02r_pytest.log
timeout 300 python profiles/kbench.py --only dwconv --stages 0,1,2,3 --iters 5 > gpurun_out/r02r_kbench_dw.jsonl 2>&1; grep -E "x3|f32" gpurun_out/r02r_kbench_dw.jsonl | head -20
timeout 300 python bench.py --no-cpu-baseline --kernels-out gpurun_out/r02r_kernels.json > gpurun_out/r02r_bench.json 2> gpurun_out/r02r_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r02r_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['roofline']['cnx_kernels_ms_per_step'], d['variants'])
k=json.load(open('gpurun_out/r02r_kernels.json'))
for f in k['families_all'][:14]: print(f['family'], f['ms'], f['bound'], f['frac'])"
timeout 300 python bench.py --no-amp --no-cpu-baseline --kernels-out gpurun_out/r02r_kernels_fp32.json > gpurun_out/r02r_bench_fp32.json 2> gpurun_out/r02r_bench_fp32.err; echo "bench fp32 rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r02r_bench_fp32.json').read().strip().splitlines()[-1])
print('fp32', d['value'], d['ms_per_step'])"
