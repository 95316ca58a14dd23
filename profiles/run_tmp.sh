mkdir -p gpurun_out
for i in 1 2 3; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29530+i)) bench.py --gpus 2 --steps 20 --warmup 5 --no-breakdown --no-variants 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e']['host_buffers_pinned'], d['e2e']['h2d_link_GBps'])"
done
