mkdir -p gpurun_out
timeout 600 python profiles/torch_gpu_step.py --amp --steps 5 2>/dev/null | tail -1
timeout 600 python profiles/torch_gpu_step.py --amp --channels-last --steps 5 2>/dev/null | tail -1
timeout 600 python profiles/torch_gpu_step.py --steps 5 2>/dev/null | tail -1
timeout 600 python bench.py --no-amp --steps 5 --warmup 3 --no-cpu-baseline --no-breakdown 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('fp32 ours', d['value'], d['ms_per_step'], d['dtype'], d['config']['workload'][:50])"
