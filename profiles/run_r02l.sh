# round 2, session l: fp32 training on the tensor cores (split-operand forward + backward GEMMs): parity suite, fp32 bench A/B,
# stock-torch GPU baselines without the per-class .item() loops
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02l_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 12 gpurun_out/r02l_pytest.log
for v in 0 1; do
  CNX_X3_TRAIN=$v timeout 400 python bench.py --no-amp --steps 8 --warmup 3 --no-cpu-baseline --no-variants --kernels-out gpurun_out/r02l_kernels_fp32_x3train$v.json > gpurun_out/r02l_bench_fp32_x3train$v.json 2> gpurun_out/r02l_bench_fp32_x3train$v.err; echo "bench fp32 x3train=$v rc=$?"
  grep '^{' gpurun_out/r02l_bench_fp32_x3train$v.json | tail -n 1 | cut -c1-200
done
timeout 300 python profiles/torch_gpu_step.py --amp --fast-loop > gpurun_out/r02l_torch_gpu_fast_amp.json 2>&1; cat gpurun_out/r02l_torch_gpu_fast_amp.json | tail -n 1
timeout 300 python profiles/torch_gpu_step.py --fast-loop > gpurun_out/r02l_torch_gpu_fast_fp32.json 2>&1; cat gpurun_out/r02l_torch_gpu_fast_fp32.json | tail -n 1
