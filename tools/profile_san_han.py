"""One forward of a small Q-SAN (2 groups x 2 blocks) and Q-HAN (10 groups x 1 block) at 8 / 4 images of 128x128 for
`ncu --profile-from-start off`: every kernel of csrc/san_han.cu appears with the launch shape of BASELINE configs[4]."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "super-resolution-meta-attention-networks_b200"))
import torch
from deepfir_b200.han_san import QHAN, QSAN
torch.manual_seed(8)
g = torch.Generator().manual_seed(3)
qsan = QSAN(n_resgroups=2, n_resblocks=2, input_para=10, scale=4, precision="bf16").cuda().eval()
qhan = QHAN(n_resgroups=10, n_resblocks=1, num_metadata=10, scale=4, precision="bf16").cuda().eval()
with torch.no_grad():
    for p in list(qsan.parameters()) + list(qhan.parameters()):
        if p.numel() == 1:
            p.normal_(0, 0.1)
    qsan.non_local.non_local.W.weight.normal_(0, 0.1)
xs = torch.rand(8, 3, 128, 128, generator=g).cuda(); ms = (torch.rand(8, 10, 1, 1, generator=g) * 0.4).cuda()
xh = torch.rand(4, 3, 128, 128, generator=g).cuda(); mh = (torch.rand(4, 10, 1, 1, generator=g) * 0.4).cuda()
with torch.no_grad():
    for _ in range(2):
        a = qsan(xs, ms); b = qhan(xh, mh)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    a = qsan(xs, ms); b = qhan(xh, mh)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
print("ok", float(a.mean()), float(b.mean()))
