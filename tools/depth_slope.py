"""Per-RCAB cost inside the real forward: Q-RCAN x4 at 32 x 128x128 with 20 / 10 / 2 blocks per group (10 groups)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "super-resolution-meta-attention-networks_b200"))
import torch
from deepfir_b200.qrcan import QRCAN


def timeit(fn, n=6, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n


B = int(os.environ.get("B", "32"))
x = torch.rand(B, 3, 128, 128, device="cuda"); meta = torch.rand(B, 10, 1, 1, device="cuda") * 0.4
res = {}
for nb in (20, 10, 2, 20):
    torch.manual_seed(8)
    net = QRCAN(n_resgroups=10, n_resblocks=nb, style="standard", num_metadata=10, include_q_layer=True, precision="bf16").cuda().eval()
    with torch.no_grad():
        res[nb] = timeit(lambda: net(x, meta))
    print("10 groups x %2d blocks: %.2f ms" % (nb, res[nb]), flush=True)
    del net
print("per RCAB: (T20-T10)/100 = %.1f us, (T10-T2)/80 = %.1f us; fixed part (T2 - 20 RCAB) = %.2f ms" %
      ((res[20] - res[10]) * 10, (res[10] - res[2]) * 12.5, res[2] - 20 * (res[20] - res[10]) / 100))
