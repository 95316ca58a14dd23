mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -n 3
timeout 600 python bench.py > gpurun_out/bench13.json 2> gpurun_out/bench13.err; echo "bench rc=$?"; tail -2 gpurun_out/bench13.err
python profiles/show_bench.py < gpurun_out/bench13.json
