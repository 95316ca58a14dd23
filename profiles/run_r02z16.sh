# round 2, session z16: ncu --set full of the FINAL fused x3 MLP kernel and wgrad pair kernel (stage 0 / stage 2 shapes)
mkdir -p gpurun_out
KB="python profiles/kbench.py --only gemm --stages 0,2 --iters 1 --warmup 1"
$KB > gpurun_out/r02z16_kb.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'mlp_fused_x3|gemm_wgrad_pair' -c 6 -o /tmp/r02z16 $KB > gpurun_out/r02z16_ncu.log 2>&1; echo "ncu rc=$?"
python profiles/ncu_summary.py /tmp/r02z16.ncu-rep gpurun_out/r02z16_ncu_full.csv
ncu -i /tmp/r02z16.ncu-rep --page source --csv > gpurun_out/r02z16_source.csv 2>/dev/null
python profiles/ncu_source_stalls.py gpurun_out/r02z16_source.csv 12 > gpurun_out/r02z16_stalls.txt; head -16 gpurun_out/r02z16_stalls.txt | cut -c1-170
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r02z16_ncu_full.csv')))
h=rows[0]
for r in rows[1:]:
    d=dict(zip(h,r))
    print(d.get('kernel','')[:60], d.get('gpu__time_duration.sum [us]'), 'tensor', d.get('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active [%]'), 'dram%', d.get('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed [%]'), 'issue', d.get('smsp__issue_active.avg.pct_of_peak_sustained_active [%]'))
PY
