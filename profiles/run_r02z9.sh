# round 2, session z9 (N GPUs): final multi-GPU bench lines of BASELINE config 2 on the end-of-round kernels; on the 8-GPU box also
# the 1-GPU line of the same box (same-box weak-scaling reference)
N=${N:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29631"
B="bench.py --no-cpu-baseline --no-variants --no-breakdown"
$TR $B --gpus $N --steps 16 --warmup 4 > gpurun_out/r02z9_n${N}_cfg2.json 2> gpurun_out/r02z9_n${N}_cfg2.err; echo "N=$N rc=$?"
if [ "$N" = "8" ]; then
  python $B --gpus 1 --steps 16 --warmup 4 > gpurun_out/r02z9_n1_on_8gpu_box_cfg2.json 2> gpurun_out/r02z9_n1_on_8gpu_box.err; echo "N=1 rc=$?"
fi
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/r02z9_n*_cfg2.json")):
    try:
        d=json.loads([l for l in open(f).read().splitlines() if l.startswith("{")][-1])
        print(f.split("/")[-1], d["n_gpus"], d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["config"]["global_batch"], d["clocks"])
    except Exception as e:
        print(f, "failed", e)
PY
