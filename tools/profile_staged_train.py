"""One training step each of a small Q-SAN (2 groups x 2 blocks), Q-HAN (10 groups x 1 block) and a pixel-attention Q-RCAN at
16 x 64x64 LR patches, for `ncu --set full` of the backward kernels of round 2 (non-local, LAM, CSAM, SOCA MLP, pixel
attention, the two generic reductions).  Launch shapes equal those of the full-depth networks (the kernels outside the trunk
do not depend on the depth)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "super-resolution-meta-attention-networks_b200"))
import torch
import torch.nn.functional as F
from deepfir_b200.han_san import QHAN, QSAN
from deepfir_b200.qrcan import QRCAN
torch.manual_seed(8)
g = torch.Generator().manual_seed(3)
B, LR = 16, 64
x = torch.rand(B, 3, LR, LR, generator=g).cuda(); y = torch.rand(B, 3, 4 * LR, 4 * LR, generator=g).cuda()
meta = (torch.rand(B, 10, 1, 1, generator=g) * 0.4).cuda()
nets = [("qsan", QSAN(n_resgroups=2, n_resblocks=2, input_para=10, scale=4, precision="bf16")),
        ("qhan", QHAN(n_resgroups=10, n_resblocks=1, num_metadata=10, scale=4, precision="bf16")),
        ("qrcan+pa", QRCAN(n_resgroups=1, n_resblocks=2, style="standard", num_metadata=10, include_q_layer=True,
                           include_pixel_attention=True, scale=4, precision="bf16"))]
for name, net in nets:
    net = net.cuda().train()
    net.cuda_graphs = False
    for p in net.parameters():          # LAM / CSAM / SAN gammas start at zero in the reference: give them a value
        if p.numel() == 1:
            p.data.fill_(0.3)
    for _ in range(2):
        net.zero_grad(set_to_none=True)
        loss = F.l1_loss(net(x, meta), y)
        loss.backward()
    torch.cuda.synchronize()
    print(name, "train step ok, loss", float(loss))
