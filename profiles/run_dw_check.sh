mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_dwconv_ln_gpu.py -x -q > gpurun_out/t_dw.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/t_dw.log
timeout 200 python profiles/kbench.py --only dwconv --iters 3 > gpurun_out/kb_dw.jsonl 2> gpurun_out/kb_dw.err; echo "kbench rc=$?"; cat gpurun_out/kb_dw.jsonl; tail -3 gpurun_out/kb_dw.err
