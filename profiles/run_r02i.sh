# round 2, session i: 256 x 384 single-accumulator tiles for the long-K N = 384 GEMMs (A/B against the 256 x 192 tiles)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_block_gpu.py -m gpu -x -q > gpurun_out/r02i_pytest_gemm.log 2>&1; echo "pytest gemm rc=$?"; tail -n 5 gpurun_out/r02i_pytest_gemm.log
for v in 0 1; do
  CNX_GEMM_BN384=$v timeout 300 python profiles/kbench.py --only gemm --stages 2,3 --iters 5 > gpurun_out/r02i_kbench_bn384_$v.jsonl 2>&1
done
python - <<'PY'
import json
rows={}
for v in (0,1):
    for l in open(f"gpurun_out/r02i_kbench_bn384_{v}.jsonl"):
        if l.startswith("{"):
            d=json.loads(l); rows.setdefault(d["kernel"],{})[v]=d["ms"]
for k,v in rows.items(): print(f"{k:32s} BN192/256: {v.get(0,0):.4f}  BN384: {v.get(1,0):.4f}  {v.get(0,1)/max(v.get(1,1),1e-9):.2f}x")
PY
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02i_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/r02i_pytest.log
timeout 300 python bench.py --no-cpu-baseline --no-variants --kernels-out gpurun_out/r02i_kernels.json > gpurun_out/r02i_bench.json 2> gpurun_out/r02i_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r02i_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['roofline']['cnx_kernels_ms_per_step'])"
