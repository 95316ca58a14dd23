# round 2, session z1: 8 vs 16 epilogue warps for the GELU-epilogue GEMMs (ring depth 5 vs 3 stages)
mkdir -p gpurun_out
for w in 0 1; do
CNX_GEMM_NEPI8=$w CNX_GEMM_STAGED=4 timeout 300 python profiles/kbench.py --only gemm --stages 0,1,2,3 --iters 5 2>&1 | grep "fc1_gelu_x3\|fc1_bias" > gpurun_out/r02z1_kbench_nepi8_$w.jsonl; echo "nepi8=$w (slab kernel forced for fc1+GELU)"; cut -c1-118 gpurun_out/r02z1_kbench_nepi8_$w.jsonl
done
