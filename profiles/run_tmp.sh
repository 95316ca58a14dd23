mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_engine_gpu.py -x -q 2>&1 | tail -n 5
N=${NGPU:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_${N}gpu.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_${N}gpu.json').read().strip().splitlines()[-1])
d.pop('kernels',None); print(d)
PY
