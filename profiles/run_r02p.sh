# round 2, session p: ncu --set full with source for the epilogue-bound GEMMs at stage 2 (fc1+GELU, x3 fc1, dGELU) — where do the
# epilogue warps wait?
mkdir -p gpurun_out
KB2="python profiles/kbench.py --only gemm --stages 2 --iters 1 --warmup 1"
$KB2 > gpurun_out/r02p_kb.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'gemm_tn_tc' -c 11 -o /tmp/r02p_gemm $KB2 > gpurun_out/r02p_ncu.log 2>&1
echo "ncu rc=$?"
ncu -i /tmp/r02p_gemm.ncu-rep --page raw --csv > gpurun_out/r02p_gemm_raw.csv 2>/dev/null
ncu -i /tmp/r02p_gemm.ncu-rep --page source --csv > gpurun_out/r02p_gemm_source.csv 2>/dev/null
ls -la /tmp/r02p_gemm.ncu-rep; du -sh gpurun_out
