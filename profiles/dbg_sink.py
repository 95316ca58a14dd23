import os, sys, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import torch.distributed as dist
import imageclassification_b200 as P
from imageclassification_b200.ddp import DistributedDataParallel
dist.init_process_group("gloo", store=dist.HashStore(), rank=0, world_size=1)
dev = torch.device("cuda")
torch.manual_seed(5)
m = P.create_model("convnext_tiny", num_classes=8, ls_init_value=1.0, drop_path_rate=0.05).to(dev)
ddp = DistributedDataParallel(m, device_ids=[0], bucket_cap_mb=4.0)
names = {p: n for n, p in m.named_parameters()}
cnt = {}
orig = ddp._ready
def ready(p):
    cnt[names[p]] = cnt.get(names[p], 0) + 1
    return orig(p)
ddp._ready = ready
x = torch.randn(4, 3, 64, 64, device=dev); t = torch.rand(4, 8, device=dev).softmax(-1)
try:
    P.SoftTargetCrossEntropy()(ddp(x), t).backward()
except Exception as e:
    print("ERR", e)
for n, p in m.named_parameters():
    if cnt.get(n, 0) != 1:
        print("count", n, cnt.get(n, 0))
print("n params", len(names), "reported", len(cnt))
