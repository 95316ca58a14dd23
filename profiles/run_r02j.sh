# round 2, session j: 16-worker 28 x 8 dwconv geometry for 56^2 maps (A/B), dwconv parity tests, bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dwconv_ln_gpu.py tests/test_block_gpu.py tests/test_full_size_properties_gpu.py -m gpu -x -q > gpurun_out/r02j_pytest_dw.log 2>&1; echo "pytest dw rc=$?"; tail -n 4 gpurun_out/r02j_pytest_dw.log
for v in 0 1; do
  CNX_DW_T28=$v timeout 300 python profiles/kbench.py --only dwconv --stages 0 --iters 5 > gpurun_out/r02j_kbench_dw_t28_$v.jsonl 2>&1
done
python - <<'PY'
import json
rows={}
for v in (0,1):
    for l in open(f"gpurun_out/r02j_kbench_dw_t28_{v}.jsonl"):
        if l.startswith("{"):
            d=json.loads(l); rows.setdefault(d["kernel"],{})[v]=d["ms"]
for k,v in rows.items(): print(f"{k:32s} 8x28 (14 workers): {v.get(0,0):.4f}  28x8 (16 workers): {v.get(1,0):.4f}  {v.get(0,1)/max(v.get(1,1),1e-9):.2f}x")
PY
timeout 300 python bench.py --no-cpu-baseline --no-variants --kernels-out gpurun_out/r02j_kernels.json > gpurun_out/r02j_bench.json 2> gpurun_out/r02j_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r02j_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['roofline']['cnx_kernels_ms_per_step'])"
