run() { echo "== $1"; env $1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus 2 --steps 20 --warmup 5 --no-breakdown --no-variants 2>/dev/null | python profiles/show_bench.py; }
run A=1 29551
run NCCL_MAX_CTAS=4 29552
run NCCL_MAX_CTAS=2 29553
run "NCCL_MAX_CTAS=8 NCCL_ALGO=Ring" 29554
