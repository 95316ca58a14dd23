# round 2, session z8: full GPU suite on the final tree (incl. the full-size fp32 accuracy forward), smoke, bench
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r02z8_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 5 gpurun_out/r02z8_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02z8_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/r02z8_smoke.log
for i in 1 2; do
timeout 300 python bench.py --no-cpu-baseline --kernels-out gpurun_out/r02z8_kernels.json > gpurun_out/r02z8_bench_$i.json 2> gpurun_out/r02z8_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r02z8_bench_$i.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['roofline']['cnx_kernels_ms_per_step'], d['variants'], d['gpu_launches'], d['clocks']['sm_mhz'])"
done
CNX_FUSED_MLP_X3=0 timeout 300 python bench.py --no-cpu-baseline --no-variants > gpurun_out/r02z8_bench_unfused.json 2> gpurun_out/r02z8_bench.err; python -c "
import json
d=json.loads(open('gpurun_out/r02z8_bench_unfused.json').read().strip().splitlines()[-1])
print('unfused x3 MLP:', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['roofline']['cnx_kernels_ms_per_step'])"
