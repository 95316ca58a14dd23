mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_engine_gpu.py -x -q 2>&1 | tail -n 25
TAG=r01g bash profiles/run_evidence.sh
