# round 2, session z10 (8 GPUs): BASELINE config 3 (ConvNeXt-Base, global batch 4096, mixup + cutmix + EMA) and config 4
# (ConvNeXt-Large 384^2, 64/GPU) on the end-of-round kernels
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29641"
B="bench.py --gpus 8 --no-cpu-baseline --no-variants --no-breakdown"
$TR $B --model convnext_base --batch 512 --update-freq 1 --cutmix 1.0 --steps 6 --warmup 3 > gpurun_out/r02z10_n8_cfg3_base.json 2> gpurun_out/r02z10_n8_cfg3.err; echo "cfg3 rc=$?"
$TR $B --model convnext_large --img 384 --batch 64 --steps 8 --warmup 3 > gpurun_out/r02z10_n8_cfg4_large384.json 2> gpurun_out/r02z10_n8_cfg4.err; echo "cfg4 rc=$?"
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/r02z10_n8_*.json")):
    try:
        d=json.loads([l for l in open(f).read().splitlines() if l.startswith("{")][-1])
        print(f.split("/")[-1], d["n_gpus"], d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["config"]["global_batch"], d["config"]["workload"][:60])
    except Exception as e:
        print(f, "failed", e)
PY
