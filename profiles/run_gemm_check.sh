mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py -x -q > gpurun_out/t_gemm.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/t_gemm.log
echo "--- default"; timeout 300 python profiles/kbench.py --only gemm --iters 3 2>&1 | grep -v wgrad
echo "--- slab16 forced"; CNX_GEMM_STAGED=4 timeout 300 python profiles/kbench.py --only gemm --iters 3 2>&1 | grep gelu
