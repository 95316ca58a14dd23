# round 2, session f: two-segment split operands for the fp32-accurate forward; full GPU suite + bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02f_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 8 gpurun_out/r02f_pytest.log
python bench.py --no-cpu-baseline --no-variants --kernels-out gpurun_out/r02f_kernels.json > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r02f_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['roofline']['cnx_kernels_ms_per_step'], d['gpu_launches'])
k=json.load(open('gpurun_out/r02f_kernels.json'))
for f in k['families_all'][:14]: print(f['family'], f['ms'], f['frac'])"
