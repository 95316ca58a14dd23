mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dwconv_ln_gpu.py tests/test_patchify_gpu.py tests/test_block_gpu.py -x -q 2>&1 | tail -n 4
for v in 1 0; do echo "== CNX_LN_V3=$v"; CNX_LN_V3=$v timeout 300 python profiles/kbench.py --only lnf --iters 3 2>&1 | grep ln_; done
