# round-1 evidence pass: launch list of the bench command + ncu --set full of the dominant kernels (one ncu tool call set)
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-breakdown --no-cpu-baseline"
$CMD > gpurun_out/ll_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 2200 -c 900 --csv --log-file gpurun_out/r01f_launches.csv $CMD > gpurun_out/ll_ncu.log 2>&1
echo "launchlist rc=$?"
KB="python profiles/kbench.py --only gemm,dwconv,ln --stages 0,2 --iters 1 --warmup 1"
$KB > gpurun_out/ncu_full_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'gemm_tn_tc|gemm_wgrad|dwconv7|ln_' -c 60 -o gpurun_out/r01f_full $KB > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"; tail -n 3 gpurun_out/ncu_full.log
