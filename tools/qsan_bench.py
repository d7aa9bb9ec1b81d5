"""BASELINE.json configs[4]: Q-SAN x4 (20 groups x 10 blocks, second-order attention + non-local) on synthetic 128x128
LR images: direct net.forward and the handler's chopped evaluation (4 quadrants of 74x74, as q-san.toml)."""
import os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "super-resolution-meta-attention-networks_b200"))
import torch
from SISR.models import ModelInterface
torch.manual_seed(8)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
h = ModelInterface.define_model("qsan", device=0, model_save_dir=tempfile.gettempdir(), eval_mode=True, scale=4,
                                metadata=["blur_kernel"], max_combined_im_size=20000)
with torch.no_grad():
    for p in (h.net.gamma, h.net.non_local.non_local.W.weight, h.net.non_local.non_local.W.bias):
        p.normal_(0, 0.1)   # the zero-initialised branches must do real work
g = torch.Generator().manual_seed(8)
x = torch.rand(B, 3, 128, 128, generator=g)
meta = torch.rand(B, 10, generator=g, dtype=torch.float64) * 0.4
keys = [("blur_kernel",) * B] * 10
xd = x.cuda(); attr = h.generate_channels(x, meta, keys).cuda()
def timeit(fn, n=3):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
with torch.no_grad():
    ms = timeit(lambda: h.net(xd, attr))
print("Q-SAN x4 direct forward, %d x 128x128: %.1f ms = %.1f output MPix/s" % (B, ms, B * 512 * 512 / 1e6 / (ms / 1e3)))
ms = timeit(lambda: h.run_eval(x, metadata=meta, metadata_keys=keys))
print("Q-SAN x4 handler run_eval (forward_chop, 4 x 74x74 quadrants, host in/out), %d images: %.1f ms = %.1f output MPix/s"
      % (B, ms, B * 512 * 512 / 1e6 / (ms / 1e3)))
