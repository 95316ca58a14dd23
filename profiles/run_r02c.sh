# round 2, session c: bench after the sync-free class counts; ncu --set full with source for the dwconv kernels (stage 0) and
# the stage-2 GEMMs (condensed to CSV on the box)
mkdir -p gpurun_out
python -m pytest tests/test_engine_gpu.py tests/test_parity_round2_gpu.py -m gpu -x -q > gpurun_out/r02c_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/r02c_pytest.log
python bench.py --no-cpu-baseline --no-variants --kernels-out gpurun_out/r02c_kernels.json > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r02c_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['roofline']['cnx_kernels_ms_per_step'])"
KB1="python profiles/kbench.py --only dwconv --stages 0 --iters 1 --warmup 1"
$KB1 > gpurun_out/r02c_kb1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'dwconv7_v2|dwconv7_wgrad_v2' -s 4 -c 4 -o /tmp/r02c_dw $KB1 > gpurun_out/r02c_ncu_dw.log 2>&1
echo "ncu dw rc=$?"
KB2="python profiles/kbench.py --only gemm --stages 2 --iters 1 --warmup 1"
$KB2 > gpurun_out/r02c_kb2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'gemm_tn_tc|gemm_wgrad' -s 11 -c 10 -o /tmp/r02c_gemm $KB2 > gpurun_out/r02c_ncu_gemm.log 2>&1
echo "ncu gemm rc=$?"
for t in dw gemm; do
  ncu -i /tmp/r02c_$t.ncu-rep --page raw --csv > gpurun_out/r02c_${t}_raw.csv 2>/dev/null
  ncu -i /tmp/r02c_$t.ncu-rep --page source --csv > gpurun_out/r02c_${t}_source.csv 2>/dev/null
  ls -la /tmp/r02c_$t.ncu-rep
done
du -sh gpurun_out
