mkdir -p gpurun_out
run() { echo "== $*"; env $1 timeout 300 python bench.py ${@:2} --no-breakdown --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['clocks'])"; }
run A=1 --steps 10 --warmup 3
run A=1 --steps 20 --warmup 5
run CNX_CLOCK_PERIOD=0.5 --steps 20 --warmup 5
run A=1 --steps 40 --warmup 5
run CNX_CLOCK_PERIOD=0.5 --steps 40 --warmup 5
