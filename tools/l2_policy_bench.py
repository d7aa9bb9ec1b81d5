"""L2 eviction-priority policies of the RCAB conv chain (DFIR_L2_POLICY, csrc/conv_tc.cu): full Q-RCAN x4 forward,
32 x 128x128, per policy string.  Letters: conv1 in, conv1 out, conv2 in, conv2 bf16 out, conv2 fp32 skip, conv2 fp32 out."""
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


torch.manual_seed(8)
B = int(os.environ.get("B", "32"))
x = torch.rand(B, 3, 128, 128, device="cuda"); meta = torch.rand(B, 10, 1, 1, device="cuda") * 0.4
net = QRCAN(n_resgroups=10, n_resblocks=20, style="standard", num_metadata=10, include_q_layer=True, precision="bf16").cuda().eval()
ref = None
pols = sys.argv[1:] or ["nnnnnn", "flflff", "nlnlff", "nlnlnn", "nnnnff", "flflnn", "nlflff", "nnnlff", "nlnnff", "nnnnnn"]
for pol in pols:
    if "=" in pol:  # any other A/B switch of the library, e.g. DFIR_WPREFETCH=0
        k, v = pol.split("=", 1); os.environ[k] = v
    else:
        os.environ["DFIR_L2_POLICY"] = pol
    with torch.no_grad():
        ms = timeit(lambda: net(x, meta)); out = net(x, meta)
    if ref is None: ref = out.clone()
    print("policy %s  %.2f ms  %.1f MPix/s  max diff vs first %.1e" % (pol, ms, B * 512 * 512 / 1e6 / (ms / 1e3), float((out - ref).abs().max())), flush=True)
