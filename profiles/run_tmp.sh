mkdir -p gpurun_out
for v in "" _U168; do echo "== variant $v"; CNX_LIB=$PWD/imageclassification_b200/lib/libcnx$v.so timeout 300 python profiles/kbench.py --only dwconv --iters 3 2>&1 | grep "ln_fwd"; done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench8.json 2> gpurun_out/bench8.err; echo "bench rc=$?"; tail -3 gpurun_out/bench8.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench8.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e'], d['clocks'], d['variants'], d['roofline'])
for k in sorted(d['kernels'], key=lambda k:-k['ms_per_step'])[:16]:
    print("  %-30s calls %5.1f ms %7.3f GB/s %s TF %s"%(k['kernel'],k['calls_per_step'],k['ms_per_step'],k['GBps'],k['TFLOPs']))
PY
