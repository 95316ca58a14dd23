mkdir -p gpurun_out
CMD="python profiles/kbench.py --only gemm --stages 0,2 --iters 1 --warmup 1"
$CMD > gpurun_out/ncu_gemm_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'gemm_tn_tc|gemm_wgrad_tc' -c 30 -o gpurun_out/r01c_gemm $CMD > gpurun_out/ncu_gemm.log 2>&1
echo "rc=$?"; tail -14 gpurun_out/ncu_gemm_plain.log; tail -3 gpurun_out/ncu_gemm.log
