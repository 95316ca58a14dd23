mkdir -p gpurun_out
export CNX_GEMM_STAGED=4
CMD="python profiles/kbench.py --only gemm --stages 0,2 --iters 1 --warmup 1"
$CMD > gpurun_out/ncu_gemm2_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'gemm_tn_tc_kernel' -c 22 -o gpurun_out/r01e_gemm_slab $CMD > gpurun_out/ncu_gemm2.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/ncu_gemm2.log
