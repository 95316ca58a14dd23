timeout 300 python -m pytest tests/test_engine_gpu.py tests/test_small_kernels_gpu.py -x -q 2>&1 | tail -n 2
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
