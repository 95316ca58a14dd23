# round-2 evidence pass (one gpurun call): bench line + reference arm, ncu launch list of the bench command, DRAM traffic per kernel
# (-> profiles/traffic.json), ncu --set full of the hot kernels, launch list of smoke()
TAG=${TAG:-r02z}
mkdir -p gpurun_out
python bench.py --kernels-out gpurun_out/${TAG}_bench_kernels_N1.json > gpurun_out/${TAG}_bench_1gpu.json 2> gpurun_out/${TAG}_bench_1gpu.err; echo "bench rc=$?"
wc -c gpurun_out/${TAG}_bench_1gpu.json
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>&1; echo "ref rc=$?"
CMD="python bench.py --steps 2 --warmup 3 --no-breakdown --no-cpu-baseline --no-variants"
$CMD > gpurun_out/${TAG}_ll_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 1200 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ll_ncu.log 2>&1
echo "launch list rc=$?"
python profiles/launch_share.py gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_launch_share.txt 2>&1; head -n 4 gpurun_out/${TAG}_launch_share.txt
KB="python profiles/kbench.py --only gemm,dwconv,ln --stages 0,2 --iters 1 --warmup 0 --call-log gpurun_out/${TAG}_kbench_calls.json"
$KB > gpurun_out/${TAG}_kbench_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/${TAG}_kbench_dram.csv $KB > gpurun_out/${TAG}_kbench_ncu.log 2>&1
echo "dram list rc=$?"
python profiles/make_traffic.py gpurun_out/${TAG}_kbench_dram.csv gpurun_out/${TAG}_kbench_calls.json gpurun_out/${TAG}_traffic.json "profiles/${TAG}_kbench_dram.csv (ncu dram__bytes_read.sum + dram__bytes_write.sum per launch of profiles/kbench.py --only gemm,dwconv,ln --stages 0,2, batch 256)"
$KB > /dev/null 2>&1 && \
ncu --set full --clock-control none -k regex:'gemm_tn_tc|gemm_wgrad|dwconv7|ln_bwd|ln_fwd' -c 48 -o /tmp/${TAG}_full $KB > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "full rc=$?"
python profiles/ncu_summary.py /tmp/${TAG}_full.ncu-rep gpurun_out/${TAG}_ncu_full.csv
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file gpurun_out/${TAG}_smoke_launches.csv python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke_ncu.log 2>&1
echo "smoke list rc=$?"; tail -n 1 gpurun_out/${TAG}_smoke.log
python profiles/launch_share.py gpurun_out/${TAG}_smoke_launches.csv > gpurun_out/${TAG}_smoke_launch_share.txt 2>&1; head -n 12 gpurun_out/${TAG}_smoke_launch_share.txt
du -sh gpurun_out
