"""One Q-RCAN forward (for ncu launch lists): python tools/fwd_once.py [B] [chunk] [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "super-resolution-meta-attention-networks_b200"))
import torch
from deepfir_b200.qrcan import QRCAN
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 8
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
LR = int(os.environ.get("LR", "128"))
torch.manual_seed(8)
net = QRCAN(n_resgroups=10, n_resblocks=20, style="standard", num_metadata=10, include_q_layer=True,
            precision="bf16", chunk_images=chunk).cuda().eval()
x = torch.rand(B, 3, LR, LR, device="cuda"); meta = torch.rand(B, 10, 1, 1, device="cuda") * 0.4
with torch.no_grad():
    for _ in range(reps):
        out = net(x, meta)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()   # ncu --profile-from-start off: list only this forward
    out = net(x, meta)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
print("ok", out.shape, float(out.abs().mean()))
