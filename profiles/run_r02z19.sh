# round 2, session z19: FINAL full GPU suite + smoke + default bench + fp32 bench (default three-launch wgrad_x3 and opt-in one-loop)
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r02z19_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r02z19_pytest.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02z19_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/r02z19_smoke.log
timeout 400 python bench.py --kernels-out gpurun_out/r02z19_bench_kernels_N1.json > gpurun_out/r02z19_bench_1gpu.json 2> gpurun_out/r02z19_bench.err; echo "bench rc=$?"; wc -c gpurun_out/r02z19_bench_1gpu.json
python -c "
import json
d=json.loads(open('gpurun_out/r02z19_bench_1gpu.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['cnx_kernels_ms_per_step'], d['variants'], d['clocks']['sm_mhz'])"
for v in 0 1; do
CNX_WGRAD_X3_ONE_LOOP=$v timeout 300 python bench.py --no-amp --no-cpu-baseline --no-variants > gpurun_out/r02z19_bench_fp32_oneloop$v.json 2> gpurun_out/r02z19_bench.err; python -c "
import json
d=json.loads(open('gpurun_out/r02z19_bench_fp32_oneloop$v.json').read().strip().splitlines()[-1])
print('fp32 one_loop=$v', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])"
done
