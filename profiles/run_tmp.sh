timeout 300 python profiles/kbench.py --only gemm --stages 0,1 --iters 3 2>&1 | grep -E "fc1|fc2_|fused"
