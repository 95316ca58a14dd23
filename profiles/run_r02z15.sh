# round 2, session z15: full GPU suite + smoke + default bench on the FINAL tree
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r02z15_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r02z15_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02z15_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/r02z15_smoke.log
timeout 400 python bench.py --kernels-out gpurun_out/r02z15_bench_kernels_N1.json > gpurun_out/r02z15_bench_1gpu.json 2> gpurun_out/r02z15_bench.err; echo "bench rc=$?"; wc -c gpurun_out/r02z15_bench_1gpu.json
python -c "
import json
d=json.loads(open('gpurun_out/r02z15_bench_1gpu.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['roof_time_frac'], d['roofline']['cnx_kernels_ms_per_step'], d['variants'], d['clocks']['sm_mhz'], d['cpu_baseline']['value'])"
