"""Where the end-to-end time of QRCANHandler.run_eval goes (32 x 128x128, pinned host input -> host output)."""
import os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "super-resolution-meta-attention-networks_b200"))
import torch
import bench
h = bench.build_net(0)
x, meta, keys = bench.synth_batch(32, 8)
xp = x.pin_memory()
def t(fn, n=5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): r = fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3, r
for _ in range(3): out = h.run_eval(xp, metadata=meta, metadata_keys=keys)[0]
ms, _ = t(lambda: h.run_eval(xp, metadata=meta, metadata_keys=keys)[0]); print("run_eval total            %.2f ms" % ms)
ms, attr = t(lambda: h.generate_channels(xp, meta, keys)); print("generate_channels (CPU)   %.2f ms" % ms)
ms, xd = t(lambda: xp.to(0)); print("H2D x                     %.2f ms" % ms)
ad = attr.to(0)
with torch.no_grad():
    ms, od = t(lambda: h.net(xd, ad)); print("forward (device)          %.2f ms" % ms)
ms, _ = t(lambda: h._to_host(od)); print("D2H via _to_host          %.2f ms  (%.1f GB/s)" % (ms, od.numel() * 4 / ms / 1e6))
buf = torch.empty(od.shape, dtype=od.dtype, pin_memory=True)
def d2h():
    buf.copy_(od, non_blocking=True); torch.cuda.current_stream().synchronize(); return buf
ms, _ = t(d2h); print("D2H into a fixed pinned buffer %.2f ms  (%.1f GB/s)" % (ms, od.numel() * 4 / ms / 1e6))
ms, _ = t(lambda: h.run_eval(xp, metadata=meta, metadata_keys=keys, keep_on_device=True)[0]); print("run_eval keep_on_device   %.2f ms" % ms)
