run() { echo "== $*"; env $* timeout 300 python bench.py --no-cpu-baseline --no-breakdown --no-variants 2>/dev/null | python profiles/show_bench.py | cut -c1-230; }
run CNX_BENCH_KEEP_GC=1 CNX_ENGINE_TUNE_GC=1
run CNX_BENCH_KEEP_GC=1 CNX_ENGINE_TUNE_GC=0
run CNX_BENCH_KEEP_GC=1 CNX_ENGINE_TUNE_GC=1
run CNX_BENCH_KEEP_GC=1 CNX_ENGINE_TUNE_GC=0
timeout 300 python -m pytest tests/test_engine_gpu.py -x -q 2>&1 | tail -n 2
