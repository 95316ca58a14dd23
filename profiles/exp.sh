mkdir -p gpurun_out
for v in "" _MINB2 _NOLN; do echo "== variant $v"; CNX_LIB=$PWD/imageclassification_b200/lib/libcnx$v.so python profiles/kbench.py --only dwconv --iters 3 2>&1 | grep -v wgrad; done > gpurun_out/exp1.log 2>&1
python profiles/kbench.py --only lnf --iters 3 >> gpurun_out/exp1.log 2>&1
cat gpurun_out/exp1.log
