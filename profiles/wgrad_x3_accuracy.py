import sys, os, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from imageclassification_b200 import _lib as L
from cabi import max_rel
lib = L.load(); DEV='cuda'; bf=torch.bfloat16; st=L.stream()
for (M,N1,N2) in [(802816,96,384),(802816,384,96),(50176,384,1536),(12544,768,3072)]:
    g = torch.Generator(device=DEV).manual_seed(M + N1)
    X = torch.randn(M, N1, device=DEV, generator=g); Y = torch.randn(M, N2, device=DEV, generator=g)
    X2 = torch.empty(M, 2*N1, dtype=bf, device=DEV); Y2 = torch.empty(M, 2*N2, dtype=bf, device=DEV)
    L.check(lib.cnx_split3(L.ptr(X), M, N1, L.ptr(X2), 2, st)); L.check(lib.cnx_split3(L.ptr(Y), M, N2, L.ptr(Y2), 2, st))
    wsb = lib.cnx_gemm_wgrad_workspace_bytes(M, N1, N2, L.dt(bf), 0); ws = torch.empty(max(wsb,16), dtype=torch.uint8, device=DEV)
    out = torch.empty(N1, N2, device=DEV); cs = torch.empty(N1, device=DEV)
    L.check(lib.cnx_gemm_wgrad_x3(L.ptr(X2), L.ptr(Y2), M, N1, N2, 0, L.ptr(out), L.ptr(cs), L.ptr(ws), wsb, st))
    ref = torch.zeros(N1, N2, dtype=torch.float64, device=DEV)
    for i in range(0, M, 65536): ref += X[i:i+65536].double().t() @ Y[i:i+65536].double()
    f32 = X.t() @ Y      # cuBLAS fp32 (context)
    print(os.environ.get('CNX_WGRAD_X3_PASSES','1'), (M,N1,N2), 'x3 err', '%.2e'%max_rel(out.double(), ref), 'torch fp32 matmul err', '%.2e'%max_rel(f32.double(), ref), 'colsum', '%.2e'%max_rel(cs.double(), X.double().sum(0)))
