"""Q-RCAN forward timing for one setting of DFIR_SPLIT (concurrent sub-passes; the library reads the variable once per
process, so run this once per setting):  DFIR_SPLIT=2 python tools/split_bench.py [B] [reps]
Prints ms per forward (CUDA events, L2 flushed between forwards) and a checksum of the output, which must not depend on
the setting (an image's result is independent of what else runs)."""
import os, sys, hashlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "super-resolution-meta-attention-networks_b200"))
import torch
from deepfir_b200.qrcan import QRCAN
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 8
LR = int(os.environ.get("LR", "128"))
torch.manual_seed(8)
net = QRCAN(n_resgroups=10, n_resblocks=20, style="standard", num_metadata=10, include_q_layer=True,
            precision="bf16").cuda().eval()
x = torch.rand(B, 3, LR, LR, device="cuda"); meta = torch.rand(B, 10, 1, 1, device="cuda") * 0.4
flush = torch.empty(192 << 20, dtype=torch.uint8, device="cuda")
with torch.no_grad():
    for _ in range(3):
        out = net(x, meta)
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = net(x, meta); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
ts.sort()
h = hashlib.sha256(out.cpu().numpy().tobytes()).hexdigest()[:16]
print("DFIR_SPLIT=%s B=%d: median %.3f ms  min %.3f ms  -> %.1f MPix/s  sha %s" % (
    os.environ.get("DFIR_SPLIT", "-"), B, ts[len(ts) // 2], ts[0], B * (4 * LR) ** 2 / 1e3 / ts[len(ts) // 2], h))
