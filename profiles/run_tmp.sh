timeout 600 python -m pytest tests/test_patchify_gpu.py -x -q 2>&1 | tail -n 15
