# round 2, session w: full GPU suite on the current tree; write/read/copy bandwidth probe; ncu --set full with source of the
# stage-0 x3 GEMMs (fc1+GELU split output, fc2) — what do they wait for at 62-76 % of HBM?
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02w_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 5 gpurun_out/r02w_pytest.log
python profiles/membw_probe.py > gpurun_out/r02w_membw.json; cat gpurun_out/r02w_membw.json
KB="python profiles/kbench.py --only gemm --stages 0 --iters 1 --warmup 1"
$KB > gpurun_out/r02w_kb.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gemm_tn_tc' -s 5 -c 4 -o /tmp/r02w_gemm $KB > gpurun_out/r02w_ncu.log 2>&1
echo "ncu rc=$?"
ncu -i /tmp/r02w_gemm.ncu-rep --page raw --csv > gpurun_out/r02w_gemm_raw.csv 2>/dev/null
ncu -i /tmp/r02w_gemm.ncu-rep --page source --csv > gpurun_out/r02w_gemm_source.csv 2>/dev/null
python profiles/ncu_source_stalls.py gpurun_out/r02w_gemm_source.csv 14 > gpurun_out/r02w_ncu_source_stalls.txt; head -70 gpurun_out/r02w_ncu_source_stalls.txt
