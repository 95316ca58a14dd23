# round 2, session g (N GPUs): multi-GPU bench lines — BASELINE config 2 (overlap on/off), config 3 (ConvNeXt-Base, global batch 4096 =
# 512/GPU x update_freq x N, mixup + cutmix + EMA), config 4 (ConvNeXt-Large 384^2, 64/GPU)
N=${N:-8}
UF=$((8 / N))
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621"
B="bench.py --gpus $N --no-cpu-baseline --no-variants --no-breakdown"
run() { tag=$1; shift; "$@" > gpurun_out/r02g_n${N}_$tag.json 2> gpurun_out/r02g_n${N}_$tag.err; echo "$tag rc=$?"; grep '^{' gpurun_out/r02g_n${N}_$tag.json | tail -n 1 | cut -c1-400; }
if [ "$N" = "8" ]; then
  NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL run cfg2_default $TR $B --steps 16 --warmup 4
  grep -E "NVLS|Channel|via|Connected|algo|Using" gpurun_out/r02g_n${N}_cfg2_default.json gpurun_out/r02g_n${N}_cfg2_default.err 2>/dev/null | grep -v '^{' | cut -c1-200 | sort | uniq -c | sort -rn | head -n 25 > gpurun_out/r02g_n${N}_nccl_info.txt
  CNX_DDP_OVERLAP=0 run cfg2_nooverlap $TR $B --steps 16 --warmup 4
  run cfg4_large384 $TR $B --model convnext_large --img 384 --batch 64 --steps 8 --warmup 3
fi
run cfg3_base $TR $B --model convnext_base --batch 512 --update-freq $UF --cutmix 1.0 --steps 6 --warmup 3
if [ "$N" != "8" ]; then run cfg2_default $TR $B --steps 16 --warmup 4; fi
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/r02g_n${N}_*.json")):
    try:
        d=json.loads([l for l in open(f).read().splitlines() if l.startswith("{")][-1])
        print(f.split("/")[-1], d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["config"]["global_batch"], d["clocks"]["reasons"])
    except Exception as e:
        print(f, "failed", e)
PY
