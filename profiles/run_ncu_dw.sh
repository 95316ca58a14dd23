mkdir -p gpurun_out
CMD="python profiles/kbench.py --only dwconv --stages 0,2 --iters 1 --warmup 1"
$CMD > gpurun_out/ncu_dw_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'dwconv7_v2_kernel' -c 8 -o gpurun_out/r01d_dw $CMD > gpurun_out/ncu_dw.log 2>&1
echo "rc=$?"; tail -8 gpurun_out/ncu_dw_plain.log; tail -3 gpurun_out/ncu_dw.log
