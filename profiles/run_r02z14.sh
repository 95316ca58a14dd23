# round 2, session z14: wgrad pair kernel stores its split-K partials through transposition slabs (coalesced rows)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_block_gpu.py tests/test_full_size_properties_gpu.py -m gpu -x -q > gpurun_out/r02z14_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r02z14_pytest.log
timeout 300 python profiles/kbench.py --only gemm --stages 1,2,3 --iters 5 2>&1 | grep wgrad > gpurun_out/r02z14_kbench_wgrad.jsonl; cut -c1-120 gpurun_out/r02z14_kbench_wgrad.jsonl
