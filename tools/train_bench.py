"""Times the Q-RCAN x4 training step (BASELINE.json configs[3]: batch 16 x 64x64 LR patches per GPU, L1, Adam) on one
GPU.  Usage: python tools/train_bench.py [steps] [precision]"""
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "super-resolution-meta-attention-networks_b200"))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    precision = sys.argv[2] if len(sys.argv) > 2 else "bf16"
    from SISR.models import ModelInterface
    torch.manual_seed(8)
    h = ModelInterface.define_model("qrcan", device=0, model_save_dir=tempfile.gettempdir(), eval_mode=False, lr=1e-4,
                                    metadata=["blur_kernel"], n_resgroups=10, n_resblocks=20, n_feats=64, scale=4,
                                    style="standard", include_q_layer=True, precision=precision,
                                    scheduler="cosine_annealing_warm_restarts",
                                    scheduler_params=dict(t_mult=1, restart_period=125000, lr_min=1e-7))
    g = torch.Generator().manual_seed(8)
    x = torch.rand(16, 3, 64, 64, generator=g).pin_memory()
    y = torch.rand(16, 3, 256, 256, generator=g).pin_memory()
    meta = torch.rand(16, 10, generator=g, dtype=torch.float64) * 0.4
    keys = [("blur_kernel",) * 16] * 10
    for _ in range(3):
        loss, _ = h.run_train(x, y, metadata=meta, metadata_keys=keys, keep_on_device=True)
    torch.cuda.synchronize()
    # phases, device time
    net = h.net
    xd, yd = x.cuda(), y.cuda()
    attr = h.generate_channels(x, meta, keys).cuda()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    tot = [0.0] * 4
    for _ in range(steps):
        ev[0].record()
        out = net(xd, attr)
        ev[1].record()
        l = h.criterion(out, yd)
        h.optimizer.zero_grad()
        ev[2].record()
        l.backward()
        ev[3].record()
        h.optimizer.step()
        ev[4].record()
        torch.cuda.synchronize()
        for i in range(4):
            tot[i] += ev[i].elapsed_time(ev[i + 1])
    print("device ms/step: forward(+repack) %.2f | loss %.2f | backward %.2f | adam %.2f | total %.2f"
          % tuple([t / steps for t in tot] + [sum(tot) / steps]))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        loss, _ = h.run_train(x, y, metadata=meta, metadata_keys=keys, keep_on_device=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    print("run_train wall ms/step %.2f  loss %.5f" % (dt * 1e3, float(loss)))


if __name__ == "__main__":
    main()
