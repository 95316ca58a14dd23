# round 2, session a: new parity tests + smoke (product first) + compact bench line + launch timeline
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 15 gpurun_out/r02a_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02a_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 3 gpurun_out/r02a_smoke.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file gpurun_out/r02a_smoke_launches.csv \
  python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02a_smoke_ncu.log 2>&1; echo "smoke ncu rc=$?"
python bench.py --kernels-out gpurun_out/r02a_bench_kernels_N1.json > gpurun_out/r02a_bench_1gpu.json 2> gpurun_out/r02a_bench_1gpu.err; echo "bench rc=$?"
wc -c gpurun_out/r02a_bench_1gpu.json; cat gpurun_out/r02a_bench_1gpu.json
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02a_bench_ref.json 2>&1; echo "ref rc=$?"; cat gpurun_out/r02a_bench_ref.json
