mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -n 15
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench7.json 2> gpurun_out/bench7.err; echo "bench rc=$?"; tail -3 gpurun_out/bench7.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench7.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e'], d['clocks'], d['variants'], d['roofline'])
for k in sorted(d['kernels'], key=lambda k:-k['ms_per_step'])[:24]:
    print("  %-30s calls %5.1f ms %7.3f GB/s %s TF %s"%(k['kernel'],k['calls_per_step'],k['ms_per_step'],k['GBps'],k['TFLOPs']))
PY
