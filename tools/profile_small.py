"""Small-depth Q-RCAN (1 group x 2 RCAB) at the bench shapes, for `ncu --set full`: one inference forward at 32 x 128x128
(kernels of the default schedule) and one training step at 16 x 64x64 (forward with stash + backward kernels).
Every kernel of the full-depth runs appears here with the same launch shape."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "super-resolution-meta-attention-networks_b200"))
import torch
import torch.nn.functional as F
from deepfir_b200.qrcan import QRCAN
torch.manual_seed(8)
kw = dict(n_resgroups=1, n_resblocks=2, style="standard", num_metadata=10, include_q_layer=True, scale=4)
g = torch.Generator().manual_seed(3)
mode = sys.argv[1] if len(sys.argv) > 1 else "both"
net = QRCAN(precision="bf16", **kw).cuda().eval()
x = torch.rand(32, 3, 128, 128, generator=g).cuda()
meta = (torch.rand(32, 10, 1, 1, generator=g) * 0.4).cuda()
if mode in ("both", "infer"):
    with torch.no_grad():
        for _ in range(2):
            out = net(x, meta)
    torch.cuda.synchronize()
    print("inference ok", float(out.mean()))
if mode == "infer":
    sys.exit(0)
net.train()
net.cuda_graphs = False
xt = torch.rand(16, 3, 64, 64, generator=g).cuda(); yt = torch.rand(16, 3, 256, 256, generator=g).cuda()
mt = (torch.rand(16, 10, 1, 1, generator=g) * 0.4).cuda()
for _ in range(2):
    net.zero_grad(set_to_none=True)
    loss = F.l1_loss(net(xt, mt), yt)
    loss.backward()
torch.cuda.synchronize()
print("train ok", float(loss))
