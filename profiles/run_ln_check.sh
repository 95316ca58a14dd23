mkdir -p gpurun_out
echo A minb3 U2 P444; CNX_LIB=$PWD/imageclassification_b200/lib/libcnx_LNA.so CNX_LN_P=444 timeout 200 python profiles/kbench.py --only ln --iters 3 --stages 0,1 2>&1 | tail -2
echo B minb4 U2 P592; CNX_LIB=$PWD/imageclassification_b200/lib/libcnx_LNB.so CNX_LN_P=592 timeout 200 python profiles/kbench.py --only ln --iters 3 --stages 0,1 2>&1 | tail -2
