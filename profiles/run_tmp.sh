mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -n 5
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench10.json 2> gpurun_out/bench10.err; echo "bench rc=$?"; tail -3 gpurun_out/bench10.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench10.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['clocks'], d['variants'], d['gpu_launches'])
PY
