mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -n 12
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench11.json 2> gpurun_out/bench11.err; echo "bench rc=$?"; tail -3 gpurun_out/bench11.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench11.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['variants'], d['gpu_launches'], d['config']['final_loss'])
PY
