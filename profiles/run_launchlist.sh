mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-breakdown --no-cpu-baseline"
$CMD > gpurun_out/ll_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 1400 --csv --log-file gpurun_out/r01e_launches.csv $CMD > gpurun_out/ll_ncu.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ll_plain.log | cut -c1-300; wc -l gpurun_out/r01e_launches.csv
