mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02m_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 25 gpurun_out/r02m_pytest.log
