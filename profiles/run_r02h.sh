# round 2, session h: is the long-K tcgen05 GEMM bound by L2 -> SM bandwidth?  Same kernels on 148 / 112 / 74 SMs.
mkdir -p gpurun_out
for sms in 148 112 74; do
  CNX_GEMM_MAXSMS=$sms python profiles/kbench.py --only gemm --stages 2,3 --iters 5 > gpurun_out/r02h_kbench_gemm_sms$sms.jsonl 2>&1
done
python - <<'PY'
import json
rows={}
for sms in (148,112,74):
    for l in open(f"gpurun_out/r02h_kbench_gemm_sms{sms}.jsonl"):
        if l.startswith("{"):
            d=json.loads(l); rows.setdefault(d["kernel"],{})[sms]=d["ms"]
print("kernel ms@148 ms@112 ms@74  (ratio 74/148; 2.0 = scales with SMs, 1.0 = bound by a shared resource)")
for k,v in rows.items():
    if 148 in v and 74 in v: print(f"{k:32s} {v[148]:.4f} {v.get(112,0):.4f} {v[74]:.4f}  {v[74]/v[148]:.2f}")
PY
python -c "import __graft_entry__ as g; g.smoke()"
