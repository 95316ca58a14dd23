mkdir -p gpurun_out
CMD="python profiles/kbench.py --only dwconv,ln --stages 0,2 --iters 1 --warmup 1"
$CMD > gpurun_out/ncu1_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'dwconv7_kernel|dwconv7_wgrad_kernel|ln_bwd' -s 8 -c 8 -o gpurun_out/r01b_bw $CMD > gpurun_out/ncu1.log 2>&1
echo "rc=$?"; tail -5 gpurun_out/ncu1_plain.log; tail -5 gpurun_out/ncu1.log
