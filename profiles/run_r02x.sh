# round 2, session x: L2 prefetch of future tiles by the GEMM producers (CNX_GEMM_PF = distance in tiles, 0 = off): A/B kbench
mkdir -p gpurun_out
for w in 0 1 2 4; do
CNX_GEMM_PF=$w timeout 300 python profiles/kbench.py --only gemm --stages 0,1,2,3 --iters 5 > gpurun_out/r02x_kbench_gemm_pf$w.jsonl 2>&1; echo "pf=$w"; python - $w <<'PY'
import json,sys
for l in open(f'gpurun_out/r02x_kbench_gemm_pf{sys.argv[1]}.jsonl'):
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print('  %-30s %.4f  hbm %.3f tensor %.3f'%(d['kernel'],d['ms'],d['hbm_frac'],d['tensor_frac']))
PY
done
