mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -n 3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 1
timeout 600 python bench.py > gpurun_out/bench14.json 2> gpurun_out/bench14.err; echo "bench rc=$?"; tail -2 gpurun_out/bench14.err
python profiles/show_bench.py < gpurun_out/bench14.json
