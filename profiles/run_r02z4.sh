# round 2, session z4: downsample conv writes the widened fp32 stream itself + takes the bf16 gradient hand-off: parity + bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02z4_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 5 gpurun_out/r02z4_pytest.log
timeout 300 python bench.py --no-cpu-baseline --kernels-out gpurun_out/r02z4_kernels.json > gpurun_out/r02z4_bench.json 2> gpurun_out/r02z4_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r02z4_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['roofline']['cnx_kernels_ms_per_step'], d['variants'], d['gpu_launches'])
k=json.load(open('gpurun_out/r02z4_kernels.json'))
t=k['timeline']; print('gap', t['gap_ms_per_step'], t['largest_gaps_ms'][:6])"
