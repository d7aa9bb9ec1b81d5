import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "super-resolution-meta-attention-networks_b200"))
import torch
from SISR.models import BaseModel
uc = lambda b: torch._C._storage_Use_Count(b.untyped_storage()._cdata)
b = torch.empty(1024, pin_memory=True); print("pinned idle", uc(b))
v = b.view(b.shape); print("with view", uc(b))
d = torch.rand(1024, device="cuda"); v.copy_(d, non_blocking=True); torch.cuda.synchronize(); print("after copy", uc(b))
del v; print("view dropped", uc(b))
od = torch.rand(32, 3, 512, 512, device="cuda")
for i in range(8):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    r = BaseModel._to_host(od)
    dt = (time.perf_counter() - t0) * 1e3
    print(i, "%.2f ms" % dt, "pool sizes", {k[0][0]: [uc(x) for x in v] for k, v in BaseModel._PINNED.items()})
