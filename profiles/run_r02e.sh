# round 2, session e (2 GPUs): NCCL parity tests, DDP modes at N=2 against N=1 on the same box
mkdir -p gpurun_out
python -m pytest tests/test_ddp_nccl_gpu.py -m gpu -x -q > gpurun_out/r02e_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 6 gpurun_out/r02e_pytest.log
B="bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-variants --no-breakdown"
python $B > gpurun_out/r02e_n1.json 2> gpurun_out/r02e_n1.err; echo "n1 rc=$?"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
$TR $B --gpus 2 > gpurun_out/r02e_n2_default.json 2> gpurun_out/r02e_n2_default.err; echo "n2 default rc=$?"
CNX_DDP_OVERLAP=0 $TR $B --gpus 2 > gpurun_out/r02e_n2_nooverlap.json 2> gpurun_out/r02e_n2_nooverlap.err; echo "n2 no-overlap rc=$?"
CNX_DDP_DIRECT=0 $TR $B --gpus 2 > gpurun_out/r02e_n2_nodirect.json 2> gpurun_out/r02e_n2_nodirect.err; echo "n2 no-direct rc=$?"
python - <<'PY'
import json
for t in ("n1","n2_default","n2_nooverlap","n2_nodirect"):
    try:
        d=json.loads([l for l in open(f"gpurun_out/r02e_{t}.json").read().strip().splitlines() if l.startswith("{")][-1])
        print(t, d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"])
    except Exception as e:
        print(t, "failed", e); print(open(f"gpurun_out/r02e_{t}.err").read()[-1500:])
PY
