# round 2, session t: staggered start of the second compute warp per scheduler in the dwconv kernels: A/B kbench
mkdir -p gpurun_out
for w in 0 300 700 1200; do
CNX_DW_STAGGER=$w timeout 300 python profiles/kbench.py --only dwconv --stages 0,1,2 --iters 5 > gpurun_out/r02t_kbench_dw_stagger$w.jsonl 2>&1; echo "stagger=$w"; python - $w <<'PY'
import json,sys
for l in open(f'gpurun_out/r02t_kbench_dw_stagger{sys.argv[1]}.jsonl'):
    try: d=json.loads(l)
    except Exception: continue
    print('  %-34s %.4f'%(d['kernel'],d['ms']))
PY
done
