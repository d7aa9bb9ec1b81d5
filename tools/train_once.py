"""One Q-RCAN x4 training step at BASELINE configs[3] (16 x 64x64 per GPU) after one warm-up step; for ncu launch lists."""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "super-resolution-meta-attention-networks_b200"))
import torch
from SISR.models import ModelInterface
torch.manual_seed(8)
h = ModelInterface.define_model("qrcan", device=0, model_save_dir=tempfile.gettempdir(), eval_mode=False, lr=1e-4,
                                metadata=["blur_kernel"], n_resgroups=10, n_resblocks=20, n_feats=64, scale=4,
                                style="standard", include_q_layer=True, precision=(sys.argv[1] if len(sys.argv) > 1 else "bf16"))
g = torch.Generator().manual_seed(8)
x = torch.rand(16, 3, 64, 64, generator=g); y = torch.rand(16, 3, 256, 256, generator=g)
meta = torch.rand(16, 10, generator=g, dtype=torch.float64) * 0.4
keys = [("blur_kernel",) * 16] * 10
h.net.cuda_graphs = False          # plain launches: every kernel shows up under its own name in the ncu list
for _ in range(2):
    loss, _ = h.run_train(x, y, metadata=meta, metadata_keys=keys, keep_on_device=True)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()   # ncu --profile-from-start off: only the third step is listed
loss, _ = h.run_train(x, y, metadata=meta, metadata_keys=keys, keep_on_device=True)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("loss", float(loss))
