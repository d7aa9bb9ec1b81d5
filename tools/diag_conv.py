"""Hardware bring-up diagnostic for the tcgen05 conv: one-hot taps + coded inputs reveal which input
(pixel, channel) lands in each output element.  Prints compact maps; run on the GPU box."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "super-resolution-meta-attention-networks_b200"))
import torch
from tests import gpu_util as G

torch.set_printoptions(linewidth=250, precision=1, sci_mode=False)
B, H, W = 1, 4, 128


def run(tap, code, desc_mode, perm=False):
    dy, dx = tap
    w = torch.zeros(64, 64, 3, 3)
    for c in range(64):
        w[c if not perm else (c + 1) % 64, c, dy, dx] = 1.0   # out ch = in ch (or +1)
    x = torch.zeros(B, 64, H, W)
    if code == "x":
        x += torch.arange(W).float().reshape(1, 1, 1, W)
    elif code == "c":
        x += torch.arange(64).float().reshape(1, 64, 1, 1)
    elif code == "y":
        x += torch.arange(H).float().reshape(1, 1, H, 1) + 1
    out, _, _ = G.conv_tc(G.nhwc_bf16(x), G.pack_bf16(w), torch.zeros(64), 0, desc_mode=desc_mode)
    return out.float().cpu()[0]  # [H][W][64]


for mode in (0, 1):
    for tap in ((1, 0), (1, 1), (1, 2), (0, 1), (2, 1)):
        print("==== desc_mode", mode, "tap(dy,dx)=", tap)
        ox = run(tap, "x", mode)
        oc = run(tap, "c", mode)
        oy = run(tap, "y", mode)
        print("x-code: out[y=1, x=0..23, c=0]  :", ox[1, :24, 0].tolist())
        print("x-code: out[y=1, x=5, c=0..15]  :", ox[1, 5, :16].tolist())
        print("x-code: out[y=1, x=120..127,c=0]:", ox[1, 120:, 0].tolist())
        print("c-code: out[y=1, x=0, c=0..63]  :", oc[1, 0, :].tolist())
        print("c-code: out[y=1, x=0..15, c=9]  :", oc[1, :16, 9].tolist())
        print("y-code: out[y=0..3, x=7, c=3]   :", oy[:, 7, 3].tolist())
        exp_x = (torch.arange(W).float() + (tap[1] - 1)).clamp(min=-1)
        okx = ((ox[1, :, 0] - exp_x).abs()[1:-1].max().item() == 0)
        okc = bool((oc[1, 5, :] == torch.arange(64).float()).all())
        print("   -> x map ok:", okx, " c map ok:", okc)
