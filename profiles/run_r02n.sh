mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02n_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 15 gpurun_out/r02n_pytest.log
timeout 400 python bench.py --no-amp --steps 8 --warmup 3 --no-cpu-baseline --no-variants --kernels-out gpurun_out/r02n_kernels_fp32.json > gpurun_out/r02n_bench_fp32.json 2> gpurun_out/r02n_bench_fp32.err; echo "bench fp32 rc=$?"
grep '^{' gpurun_out/r02n_bench_fp32.json | tail -n 1 | cut -c1-220
timeout 300 python bench.py --no-cpu-baseline --no-variants --no-breakdown > gpurun_out/r02n_bench.json 2> gpurun_out/r02n_bench.err; echo "bench rc=$?"
grep '^{' gpurun_out/r02n_bench.json | tail -n 1 | cut -c1-220
