mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_block_gpu.py -x -q > gpurun_out/t_gemm.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/t_gemm.log
timeout 300 python profiles/kbench.py --only gemm --iters 3 > gpurun_out/kb_gemm.jsonl 2> gpurun_out/kb_gemm.err; echo "kbench rc=$?"; grep -E "gelu" gpurun_out/kb_gemm.jsonl; tail -3 gpurun_out/kb_gemm.err
echo "--- staged"
CNX_GEMM_STAGED=3 timeout 300 python profiles/kbench.py --only gemm --iters 3 2>&1 | grep gelu
