# round 2, session s: 28-worker tile geometries for the dwconv backward-data kernel: parity + A/B kbench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dwconv_ln_gpu.py -m gpu -x -q > gpurun_out/r02s_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/r02s_pytest.log
for w in 0 1; do
CNX_DW_WIDE=$w timeout 300 python profiles/kbench.py --only dwconv --stages 0,1,2,3 --iters 5 > gpurun_out/r02s_kbench_dw_wide$w.jsonl 2>&1; echo "wide=$w"; grep -E "dgrad" gpurun_out/r02s_kbench_dw_wide$w.jsonl | head -20
done
