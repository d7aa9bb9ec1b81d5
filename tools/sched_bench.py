import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "super-resolution-meta-attention-networks_b200"))
import torch
from deepfir_b200.qrcan import QRCAN
def timeit(fn, n=5, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
torch.manual_seed(8)
x = torch.rand(32, 3, 128, 128, device="cuda"); meta = torch.rand(32, 10, 1, 1, device="cuda") * 0.4
outs = {}
for sched in ("linear", "linear3"):
    torch.manual_seed(8)
    net = QRCAN(n_resgroups=10, n_resblocks=20, style="standard", num_metadata=10, include_q_layer=True, precision="bf16", schedule=sched).cuda().eval()
    with torch.no_grad():
        ms = timeit(lambda: net(x, meta)); outs[sched] = net(x, meta)
    print("schedule %-8s %.2f ms  %.1f MPix/s" % (sched, ms, 32 * 512 * 512 / 1e6 / (ms / 1e3)))
print("max diff", float((outs["linear"] - outs["linear3"]).abs().max()))
