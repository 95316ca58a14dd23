timeout 900 python -m pytest tests/test_full_size_properties_gpu.py -x -q 2>&1 | tail -n 25
