"""Times one training step (forward + L1 + backward + Adam) of Q-SAN (20 groups x 10 blocks) and Q-HAN (10 x 20) at
BASELINE.json configs[3]'s batch shape (16 x 64x64 LR patches, x4) through the handlers' run_train, device time.
Usage: python tools/staged_train_bench.py [steps]"""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "super-resolution-meta-attention-networks_b200"))
import torch
from SISR.models import ModelInterface
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
g = torch.Generator().manual_seed(8)
x = torch.rand(16, 3, 64, 64, generator=g).pin_memory(); y = torch.rand(16, 3, 256, 256, generator=g).pin_memory()
meta = torch.rand(16, 10, generator=g, dtype=torch.float64) * 0.4
keys = [("blur_kernel",) * 16] * 10
for model in ("qsan", "qhan"):
    for precision in ("bf16", "fp32"):
        torch.manual_seed(8)
        h = ModelInterface.define_model(model, device=0, model_save_dir=tempfile.gettempdir(), eval_mode=False, lr=1e-4,
                                        metadata=["blur_kernel"], scale=4, precision=precision)
        losses = []
        for _ in range(2):
            losses.append(float(h.run_train(x, y, metadata=meta, metadata_keys=keys)[0]))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            losses.append(float(h.run_train(x, y, metadata=meta, metadata_keys=keys)[0]))
        e1.record(); torch.cuda.synchronize()
        print("%s %s: %.1f ms per run_train step (16 x 64x64), loss %.4f -> %.4f, peak memory %.1f GiB"
              % (model, precision, e0.elapsed_time(e1) / steps, losses[0], losses[-1], torch.cuda.max_memory_allocated() / 2 ** 30))
        del h
        torch.cuda.empty_cache()
