"""Debug helper: multi-row-per-CTA conv cases + pipeline watchdog read-out (run on the GPU box)."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "super-resolution-meta-attention-networks_b200"))
import torch
import torch.nn.functional as F
from tests import gpu_util as G
lib = G.lib()
def wd(tag):
    torch.cuda.synchronize()
    out = (C.c_uint * 8)()
    lib.dfir_debug_watchdog(C.byref(out), 2); lib.dfir_debug_watchdog(C.byref(out), 1)
    print(tag, "watchdog:", list(out))
for (B, H, W) in [(3, 160, 128), (8, 128, 128), (40, 9, 128), (3, 50, 300)]:
    g = torch.Generator().manual_seed(0)
    x = torch.rand(B, 64, H, W, generator=g) - 0.5
    w = (torch.rand(64, 64, 3, 3, generator=g) - 0.5) / 12
    b = torch.rand(64, generator=g)
    out, _, _ = G.conv_tc(G.nhwc_bf16(x), G.pack_bf16(w), b, 0)
    wd("shape %s rows/CTA %.2f" % ((B, H, W), B * H / 148.0))
    ref = F.conv2d(G.bf16_round(x), G.bf16_round(w), b, padding=1)
    print("   max err", (G.to_nchw(out) - ref).abs().max().item())
