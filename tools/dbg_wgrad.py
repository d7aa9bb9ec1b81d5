"""Bring-up diagnostic for the tcgen05 weight-gradient kernel: per-tap relative error against fp64 autograd."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "super-resolution-meta-attention-networks_b200"))
import torch, torch.nn.functional as F
from deepfir_b200 import _lib
lib = _lib.load_library()
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
def run(B, H, W):
    g = torch.Generator().manual_seed(2)
    x = torch.randn(B, 64, H, W, generator=g).to(torch.bfloat16).float()
    dy = torch.randn(B, 64, H, W, generator=g).to(torch.bfloat16).float()
    w = torch.zeros(64, 64, 3, 3, dtype=torch.float64, requires_grad=True)
    b = torch.zeros(64, dtype=torch.float64, requires_grad=True)
    F.conv2d(x.double(), w, b, padding=1).backward(dy.double())
    n = lib.dfir_conv3x3_wgrad_scratch_bytes(B, H, W, 64, 64, 0)
    scratch = torch.zeros(n, dtype=torch.uint8, device="cuda")
    dw = torch.full((64, 64, 3, 3), float("nan"), device="cuda"); db = torch.full((64,), float("nan"), device="cuda")
    xd = x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).cuda(); dyd = dy.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).cuda()
    rc = lib.dfir_conv3x3_wgrad_c64(dyd.data_ptr(), 0, 0, 0, xd.data_ptr(), B, H, W, dw.data_ptr(), db.data_ptr(), 0, 1,
                                    scratch.data_ptr(), n, st())
    torch.cuda.synchronize()
    wd = (C.c_uint * 8)()
    lib.dfir_debug_watchdog(C.byref(wd), 1)
    ref = w.grad.float()
    out = dw.cpu()
    errs = [[float((out[:, :, i, j] - ref[:, :, i, j]).norm() / ref[:, :, i, j].norm()) for j in range(3)] for i in range(3)]
    print("B,H,W", (B, H, W), "rc", rc, "watchdog", list(wd), "db err", float((db.cpu() - b.grad.float()).norm() / b.grad.float().norm()))
    for r in errs:
        print("   tap rel err", ["%.2e" % e for e in r])
for shp in [(1, 1, 16), (1, 4, 16), (2, 7, 9), (1, 5, 128), (2, 3, 150), (16, 64, 64)]:
    run(*shp)
