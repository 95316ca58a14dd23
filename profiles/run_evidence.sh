# round-1 evidence pass: launch list of the bench command + ncu --set full of the dominant kernels.
# The .ncu-rep files are condensed to CSV ON THE BOX (gpurun only brings back <= 64 MiB); one small rep with sources is kept.
TAG=${TAG:-r01i}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-breakdown --no-cpu-baseline"
$CMD > gpurun_out/ll_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 2600 -c 1000 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/ll_ncu.log 2>&1
echo "launchlist rc=$?"
KB="python profiles/kbench.py --only gemm,dwconv,ln --stages 0,2 --iters 1 --warmup 0"
$KB > gpurun_out/ncu_full_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'gemm_tn_tc|gemm_wgrad|mlp_fused|dwconv7|ln_' -c 60 -o /tmp/${TAG}_full $KB > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"; tail -n 3 gpurun_out/ncu_full.log
python profiles/ncu_summary.py /tmp/${TAG}_full.ncu-rep gpurun_out/${TAG}_ncu_full.csv
ncu -i /tmp/${TAG}_full.ncu-rep --page details --csv > gpurun_out/${TAG}_ncu_details.csv 2>/dev/null
ls -la /tmp/${TAG}_full.ncu-rep
# keep the report itself only when it fits
[ $(stat -c %s /tmp/${TAG}_full.ncu-rep) -lt 50000000 ] && cp /tmp/${TAG}_full.ncu-rep gpurun_out/
du -sh gpurun_out
