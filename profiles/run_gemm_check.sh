mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py -x -q > gpurun_out/t_gemm.log 2>&1; echo "pytest rc=$?"; tail -n 12 gpurun_out/t_gemm.log
echo "--- pair"; timeout 300 python profiles/kbench.py --only gemm --iters 3 2>&1 | grep wgrad
