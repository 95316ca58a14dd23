# round 2, session z6: ncu --set full with source of the fused x3 MLP kernel (C = 96, 56^2): where does a chunk's time go?
mkdir -p gpurun_out

KB="python profiles/kbench.py --only gemm --stages 0 --iters 1 --warmup 1"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'mlp_fused_x3' -c 2 -o /tmp/r02z6 $KB > gpurun_out/r02z6_ncu.log 2>&1
echo "ncu rc=$?"
ncu -i /tmp/r02z6.ncu-rep --page raw --csv > gpurun_out/r02z6_raw.csv 2>/dev/null
ncu -i /tmp/r02z6.ncu-rep --page source --csv > gpurun_out/r02z6_source.csv 2>/dev/null
python profiles/ncu_source_stalls.py gpurun_out/r02z6_source.csv 24 > gpurun_out/r02z6_stalls.txt; head -30 gpurun_out/r02z6_stalls.txt
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r02z6_raw.csv')))
h=rows[0]
want=['gpu__time_duration.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__inst_executed.sum','lts__t_sector_hit_rate.pct']
r=rows[2] if len(rows)>2 else rows[1]
for w in want:
    for i,x in enumerate(h):
        if x==w: print(w, r[i])
PY
