# 1-box scaling check: bench.py under torchrun at the N given in NLIST (weak scaling, batch 256/GPU)
mkdir -p gpurun_out
for N in ${NLIST:-4 8}; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29520+N)) bench.py --gpus $N --steps 20 --warmup 5 --no-breakdown > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "N=$N bench rc=$?"; tail -2 gpurun_out/bench_${N}gpu.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_${N}gpu.json').read().strip().splitlines()[-1])
print($N, d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['clocks'])
PY
done
