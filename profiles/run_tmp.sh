for v in 0 1; do echo "== CNX_GEMM_NCTA=$v"; CNX_GEMM_NCTA=$v timeout 300 python profiles/kbench.py --only gemm --stages 2,3 --iters 5 2>&1 | grep -E "fc1|fc2|dgrad|wgrad" | cut -c1-110; done
