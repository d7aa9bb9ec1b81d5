"""2-rank NCCL check of data-parallel training on the B200 path: each rank trains on its own half of a batch; after one
backward the (all-reduced) gradients must equal the single-process gradients of the whole batch.
Launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/ddp_check.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "super-resolution-meta-attention-networks_b200"))
import torch, torch.distributed as dist
import torch.nn.functional as F
from deepfir_b200.qrcan import QRCAN
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.manual_seed(8)
kw = dict(n_resgroups=2, n_resblocks=2, style="standard", num_metadata=10, include_q_layer=True, scale=4)
g = torch.Generator().manual_seed(3)
B = 2 * world
x = torch.rand(B, 3, 24, 20, generator=g).cuda(); y = torch.rand(B, 3, 96, 80, generator=g).cuda()
meta = (torch.rand(B, 10, 1, 1, generator=g) * 0.4).cuda()
for precision in ("fp32", "bf16"):
    torch.manual_seed(8)
    net = QRCAN(precision=precision, **kw).cuda().train()
    net.ddp_allreduce = False
    F.l1_loss(net(x, meta), y).backward()                     # whole batch, no exchange
    full = {k: p.grad.clone() for k, p in net.named_parameters()}
    net.zero_grad(set_to_none=True)
    net.ddp_allreduce = True
    a, b = rank * 2, rank * 2 + 2
    F.l1_loss(net(x[a:b].contiguous(), meta[a:b].contiguous()), y[a:b].contiguous()).backward()   # this rank's half, all-reduced
    worst = max(float((p.grad - full[k]).norm() / full[k].norm().clamp_min(1e-12)) for k, p in net.named_parameters())
    tol = 1e-4 if precision == "fp32" else 3e-2
    print("rank %d %s: worst relative difference of all-reduced gradients vs whole-batch gradients %.3e" % (rank, precision, worst))
    assert worst < tol, worst
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
