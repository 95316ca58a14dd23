mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/tests.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/tests.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench3.json 2> gpurun_out/bench3.err; echo "bench rc=$?"; tail -3 gpurun_out/bench3.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench3.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e'], d['roofline'])
for k in sorted(d['kernels'], key=lambda k:-k['ms_per_step'])[:45]:
    print("  %-30s calls %5.1f ms %7.3f GB/s %s TF %s"%(k['kernel'],k['calls_per_step'],k['ms_per_step'],k['GBps'],k['TFLOPs']))
print(d.get('cpu_baseline'))
PY
