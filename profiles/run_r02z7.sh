# round 2, session z7: fused x3 MLP after hoisting the bias loads and prefetching residual / next xn tile into L2
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_block_gpu.py -m gpu -x -q -k "fused_x3" > gpurun_out/r02z7_pytest_a.log 2>&1; echo "pytest(fused x3) rc=$?"; tail -n 3 gpurun_out/r02z7_pytest_a.log
timeout 300 python profiles/kbench.py --only gemm --stages 0 --iters 5 2>&1 | grep "x3" > gpurun_out/r02z7_kbench.jsonl; cut -c1-120 gpurun_out/r02z7_kbench.jsonl
