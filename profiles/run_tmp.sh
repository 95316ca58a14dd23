mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dwconv_ln_gpu.py tests/test_block_gpu.py -x -q 2>&1 | tail -n 4
timeout 300 python profiles/kbench.py --only dwconv --iters 3 2>&1 | grep "ln_fwd"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench9.json 2> gpurun_out/bench9.err; echo "bench rc=$?"; tail -3 gpurun_out/bench9.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench9.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e'], d['clocks'], d['variants'], d['roofline'])
PY
