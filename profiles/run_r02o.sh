# round 2, session o: compute-sanitizer memcheck of one small training step (smoke) — ordinary launches (CNX_PDL=0) so that the tool sees
# plain stream-ordered kernels; then the engine tests
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02o_smoke_plain.log 2>&1 && \
CNX_PDL=0 timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02o_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -n 6 gpurun_out/r02o_memcheck.log
timeout 600 python -m pytest tests/test_engine_gpu.py -m gpu -q > gpurun_out/r02o_pytest_engine.log 2>&1; echo "pytest rc=$?"; tail -n 5 gpurun_out/r02o_pytest_engine.log
