#!/usr/bin/env python
"""Deep-FIR hot-path benchmark (contract: see the task's bench.py section / DESIGN.md §Measurement).

Workload (BASELINE.json configs[1]): Q-RCAN x4 (10 groups x 20 RCAB, 64 ch, 10-D blur-kernel metadata),
batched inference on synthetic 128x128 LR images, 32 images per GPU (256 over 8 GPUs; weak scaling:
every rank runs the same per-GPU batch, no data-path collective).  One "step" = one forward pass over
the per-GPU batch.  metric = output megapixels per second, whole job.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

`--impl reference` times the CPU restatement of the reference's own path (oracle port, torch-CPU fp32 on
all host cores) on a bounded sample of the same workload; rank 0 only.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "super-resolution-meta-attention-networks_b200"))

import torch  # noqa: E402

IMAGES_PER_GPU = 32
LR = 128
SCALE = 4
NET_KW = dict(n_resgroups=10, n_resblocks=20, n_feats=64, scale=SCALE, style="standard", include_q_layer=True)
FLOP_PER_LR_PIXEL = 31835520            # conv FLOPs of Q-RCAN x4 per LR pixel
CONV64_FLOP_PER_PIXEL = 2 * 64 * 64 * 9  # one 64->64 3x3 conv


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as fh:
            d = json.load(fh)
        return dict(hbm_gbs=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"],
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region.

    nvidia-smi needs 0.1-0.3 s before its first line, longer than a short timed region, so the sampler is started
    ahead of the warm-up steps and only the lines whose timestamp falls inside [mark_begin, mark_end] are used."""

    FIELDS = ("timestamp,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.path = None
        self.t0 = self.t1 = None
        self.t1_fallback = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def mark_begin(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def mark_fallback_end(self):
        self.t1_fallback = time.time()

    @staticmethod
    def _stamp(text):
        import datetime
        return datetime.datetime.strptime(text, "%Y/%m/%d %H:%M:%S.%f").timestamp()

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"], samples=0)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        rows = []
        with open(self.path) as fh:
            for line in fh:
                parts = [t.strip() for t in line.split(",")]
                if len(parts) < 7:
                    continue
                try:
                    rows.append((self._stamp(parts[0]), float(parts[1]), float(parts[2]), parts[3:7]))
                except ValueError:
                    continue
        os.unlink(self.path)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

        def summarise(lo, hi, window):
            sel = [r for r in rows if lo <= r[0] <= hi]
            if not sel:
                return None
            sm = sorted(r[1] for r in sel)
            reasons = set()
            for r in sel:
                for n, v in zip(names, r[3]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            return dict(sm_mhz=sm[len(sm) // 2], sm_max_mhz=sel[-1][2], reasons=sorted(reasons), samples=len(sm),
                        window=window)

        out = None
        if self.t0 is not None and self.t1 is not None:
            out = summarise(self.t0, self.t1, "device-timed steps")
            if out is None and self.t1_fallback is not None:
                out = summarise(self.t0, self.t1_fallback, "device-timed steps + end-to-end steps (same load)")
        if out is None:
            out = dict(sm_mhz=None, sm_max_mhz=None, reasons=["no nvidia-smi sample inside the timed region"], samples=0)
        return out


def build_net(device):
    from SISR.models import ModelInterface
    torch.manual_seed(8)
    handler = ModelInterface.define_model(
        "qrcan", device=device, model_save_dir=tempfile.gettempdir(), eval_mode=True, lr=1e-4,
        metadata=["blur_kernel"], precision="bf16", **NET_KW)
    return handler


TRAIN_BATCH = 16                        # BASELINE.json configs[3]: 16 x 64x64 LR patches per GPU
TRAIN_LR = 64
TRAIN_FLOP_PER_STEP = 3 * FLOP_PER_LR_PIXEL * TRAIN_BATCH * TRAIN_LR * TRAIN_LR  # fwd + dgrad + wgrad


def train_step_bench(local, dev, world, dist, steps, warmup):
    """Q-RCAN x4 bf16 training step (configs[3]) through `QRCANHandler.run_train`: H2D of the LR/HR patches from pinned
    memory, forward, L1 loss, backward (all parameter gradients), gradient all-reduce over NCCL when world > 1, Adam,
    cosine scheduler, D2H of the loss.  Returns per-step times: wall clock of the API call (max over ranks) and the
    device time of forward+loss+backward+Adam on device-resident patches (CUDA events)."""
    from SISR.models import ModelInterface
    torch.manual_seed(8)
    h = ModelInterface.define_model(
        "qrcan", device=local, model_save_dir=tempfile.gettempdir(), eval_mode=False, lr=1e-4,
        metadata=["blur_kernel"], precision="bf16", scheduler="cosine_annealing_warm_restarts",
        scheduler_params=dict(t_mult=1, restart_period=125000, lr_min=1e-7), **NET_KW)
    g = torch.Generator().manual_seed(88 + int(os.environ.get("RANK", "0")))
    x = torch.rand(TRAIN_BATCH, 3, TRAIN_LR, TRAIN_LR, generator=g).pin_memory()
    y = torch.rand(TRAIN_BATCH, 3, TRAIN_LR * SCALE, TRAIN_LR * SCALE, generator=g).pin_memory()
    meta = torch.rand(TRAIN_BATCH, 10, generator=g, dtype=torch.float64) * 0.4
    keys = [("blur_kernel",) * TRAIN_BATCH] * 10

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    losses = []
    for _ in range(max(warmup, 3)):
        loss, _ = h.run_train(x, y, metadata=meta, metadata_keys=keys, keep_on_device=True)
        losses.append(float(loss))
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        loss, _ = h.run_train(x, y, metadata=meta, metadata_keys=keys, keep_on_device=True)
        losses.append(float(loss))
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3 / steps
    # device time with the patches already in HBM
    xd, yd = x.to(dev), y.to(dev)
    attr = h.generate_channels(x, meta, keys).to(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(steps):
        out = h.net(xd, attr)
        l = h.criterion(out, yd)
        h.optimizer.zero_grad()
        l.backward()
        h.optimizer.step()
    e1.record()
    barrier()
    dev_ms = e0.elapsed_time(e1) / steps
    t = torch.tensor([wall_ms, dev_ms], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    wall_ms, dev_ms = t.tolist()
    import ctypes as C
    from deepfir_b200 import _lib
    launches = int(_lib.load_library().dfir_qrcan_train_launch_count(C.byref(h.net.packed(training=True).desc),
                                                                      TRAIN_BATCH, TRAIN_LR, TRAIN_LR, 0))
    pk = peaks()
    res = {"workload": "Q-RCAN x4 (10x20 RCAB, 64 ch) bf16 training step, batch %d x %dx%d LR patches per GPU, L1, "
                       "Adam, cosine-restart scheduler%s" % (TRAIN_BATCH, TRAIN_LR, TRAIN_LR,
                                                             ", gradient all-reduce over NCCL" if world > 1 else ""),
           "ms_per_step": round(wall_ms, 3), "ms_per_step_device_resident": round(dev_ms, 3), "unit": "ms",
           "api": "QRCANHandler.run_train(x_pinned, y_pinned, metadata=, metadata_keys=, keep_on_device=True)",
           "h2d_bytes_per_step": int(x.numel() * 4 + y.numel() * 4 + TRAIN_BATCH * 10 * 4), "d2h_bytes_per_step": 4,
           "gpu_launches_per_step": launches, "cuda_graphs": True,
           "tflops_algorithmic": round(TRAIN_FLOP_PER_STEP / (dev_ms / 1e3) / 1e12, 1),
           "frac_of_bf16_peak": round(TRAIN_FLOP_PER_STEP / (dev_ms / 1e3) / 1e12 / pk["tf_sustained"], 4),
           "loss_first_last": [round(losses[0], 5), round(losses[-1], 5)]}
    del h
    torch.cuda.empty_cache()
    return res


def cpu_train_baseline(threads=None):
    """one training step of the reference's path on the host CPU: autograd through the oracle port (forward, L1,
    backward) on ONE 64x64 patch of the 16-patch batch"""
    from oracle import deepfir_oracle as O
    from deepfir_b200.qrcan import QRCAN
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    torch.manual_seed(8)
    net = QRCAN(num_metadata=10, **NET_KW)
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in net.state_dict().items()}
    g = torch.Generator().manual_seed(8)
    x = torch.rand(1, 3, TRAIN_LR, TRAIN_LR, generator=g)
    y = torch.rand(1, 3, TRAIN_LR * SCALE, TRAIN_LR * SCALE, generator=g)
    attr = (torch.rand(1, 10, generator=g) * 0.4).reshape(1, 10, 1, 1)
    t0 = time.perf_counter()
    out = O.qrcan_forward(x, attr, sd, style="standard")
    torch.nn.functional.l1_loss(out, y).backward()
    dt = time.perf_counter() - t0
    return {"value": round(dt * 1e3, 1), "unit": "ms", "cores": threads, "kind": "port",
            "sample": "forward + L1 + backward of 1 of the %d patches of a step (no warm-up, no optimizer), fp32 torch-CPU "
                      "autograd through oracle/deepfir_oracle.py; a full step is ~%dx this" % (TRAIN_BATCH, TRAIN_BATCH)}


def workload_config(n_img=None):
    n_img = n_img or IMAGES_PER_GPU
    return {"workload": "Q-RCAN x4 (10x20 RCAB, 64 ch, 10-D blur metadata) batched inference, "
                        "%d synthetic 128x128 LR images per GPU" % n_img,
            "images_per_gpu": n_img, "lr_size": [LR, LR], "scale": SCALE, "precision": "bf16 operands, fp32 accumulate, "
            "residual stream = 24-bit floats as a bf16 hi plane + an 8-bit lo plane (16 significant bits)", "l2": "192 MiB buffer rewritten between timed steps",
            "sharding": "by image, no data-path collective"}


def synth_batch(n, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.floor(torch.rand(n, 3, LR, LR, generator=g) * 256) / 255.0
    meta = torch.rand(n, 10, generator=g, dtype=torch.float64) * 0.4
    keys = [("blur_kernel",) * n] * 10
    return x, meta, keys


def flush_l2(buf):
    buf.add_(1)


def run_ours(args):
    from deepfir_b200 import _lib
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    # rank 0 times the CPU baselines BEFORE the process group exists: the other ranks then wait in the rendezvous (idle),
    # not in an NCCL barrier that keeps their GPUs spinning
    cpu = cpu_train = None
    if rank == 0:
        cpu = cpu_baseline(sample_images=2)
        cpu_train = cpu_train_baseline()
    if world > 1:
        import datetime
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(minutes=30))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    lib = _lib.load_library()
    _lib.check(lib.dfir_check_device(), "device check")

    handler = build_net(local)
    net = handler.net
    n_img = IMAGES_PER_GPU
    x_host, meta, keys = synth_batch(n_img, seed=8 + rank)
    x_pin = x_host.pin_memory()
    x_dev = x_host.to(dev)
    attr_dev = handler.generate_channels(x_host, meta, keys).to(dev)
    l2buf = torch.zeros(192 * 1024 * 1024 // 4, device=dev)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        with torch.no_grad():
            return net(x_dev, attr_dev)

    def step_e2e():
        out, _, _ = handler.run_eval(x_pin, metadata=meta, metadata_keys=keys)  # H2D + forward + D2H
        return out

    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        out = step_device()
    torch.cuda.synchronize()

    # ---------------- device-resident timing: K steps, L2 flushed between steps (flush not timed)
    barrier()
    sampler.mark_begin()
    evs = []
    wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush_l2(l2buf)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = step_device()
        e1.record()
        evs.append((e0, e1))
    barrier()
    wall = time.perf_counter() - wall0
    sampler.mark_end()
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)

    # ---------------- end to end through the handler API with host buffers
    for _ in range(3):  # the caching host allocator needs two live result buffers before the loop is steady
        out_host = step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out_host = step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0
    sampler.mark_fallback_end()
    clocks = sampler.stop()

    d2h_bytes = out_host.numel() * 4
    t = torch.tensor([dev_ms, e2e_s * 1e3], device=dev, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = t.tolist()

    # ---------------- roofline of the dominant kernel (trunk 64->64 conv), timed alone with CUDA events
    roof = roof_conv1 = None
    if rank == 0:
        roof, roof_conv1 = conv_roofline(lib, dev, net)
    # ---------------- training step (BASELINE.json metric part 2: "train step ms")
    del out, out_host, l2buf
    torch.cuda.empty_cache()
    train = train_step_bench(local, dev, world, dist, steps=max(args.steps, 5), warmup=args.warmup)
    out_mpix_step = world * n_img * (LR * SCALE) ** 2 / 1e6
    launches = int(lib.dfir_qrcan_launch_count(__import__("ctypes").byref(net.packed().desc), n_img, LR, LR, 0))

    if rank == 0:
        pk = peaks()
        ms_step = dev_ms / args.steps
        value = out_mpix_step / (ms_step / 1e3)
        extra = other_configs(dev) if world == 1 else {}
        line = {
            "metric": "Q-RCAN x4 output MPix/s", "value": round(value, 3), "unit": "MPix/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": round(ms_step, 4),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(n_img),
            "clocks": clocks,
            "e2e": {"value": round(out_mpix_step / (e2e_ms / 1e3 / args.steps), 3), "unit": "MPix/s",
                    "h2d_bytes_per_step": int(x_pin.numel() * 4 + n_img * 10 * 4),
                    "d2h_bytes_per_step": int(d2h_bytes),
                    "api": "QRCANHandler.run_eval(x_pinned_host, metadata=, metadata_keys=) -> host tensor"},
            "gpu_launches": launches * args.steps,
            "roofline": roof,
            "roofline_conv1": roof_conv1,
            "cpu_baseline": cpu,
            "train": dict(train, cpu_baseline=cpu_train),
            "tflops_conv_algorithmic": round(world * n_img * LR * LR * FLOP_PER_LR_PIXEL / (ms_step / 1e3) / 1e12, 2),
            "wall_s_timed_region": round(wall, 3),
            "peaks": pk["source"],
        }
        line.update(extra)
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


# ncu evidence for the two trunk kernels at the bench shape (32 images per launch), committed under profiles/: DRAM bytes per
# launch from the `--set full` captures, in-step duration and share of the step from the launch list of one forward
NCU = {"source": "profiles/r02_tc_kernels_ncu.md, profiles/r02_launches_infer_32x128.csv",
       "conv2_traffic": None, "conv2_in_step_us": None, "conv2_share": None, "conv1_traffic": None, "conv1_in_step_us": None,
       "conv1_share": None}
_ncu_path = next((p for p in (os.path.join(ROOT, "profiles", n) for n in ("r02c_ncu_numbers.json", "r02b_ncu_numbers.json",
                                                                         "r02_ncu_numbers.json")) if os.path.isfile(p)), "")
if os.path.isfile(_ncu_path):
    with open(_ncu_path) as _fh:
        NCU.update(json.load(_fh))


def conv_roofline(lib, dev, net):
    """Times the two trunk kernels alone (CUDA events on the launching stream), each in the form the default schedule
    launches it at 32 images x 128x128x64:
      conv1 = conv3x3 + bias + ReLU + the statistics of pool-by-linearity (dfir_conv3x3_c64_stats): tensor bound,
      conv2 = conv3x3 + in-kernel channel/meta attention + scale + residual on the hi / 8-bit lo stream, descending
              traversal (dfir_conv3x3_c64_scale_skip_hl8): 8 algorithmic bytes per element against the HBM peak."""
    import ctypes as C
    from deepfir_b200 import _lib
    pk = peaks()
    bc = IMAGES_PER_GPU
    wp = torch.empty(9 * 64 * 128, dtype=torch.uint8, device=dev)
    w = (torch.randn(64, 64, 3, 3, device=dev) / 48).contiguous()
    bias = torch.zeros(64, device=dev)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.dfir_pack_conv3x3_bf16(w.data_ptr(), wp.data_ptr(), 64, 64, 64, 0, 1, st), "pack")
    NB = 3  # rotate over three buffer sets (3 x 200 MB > L2): every launch streams from HBM like in the real chain
    xh = [torch.randn(bc, LR, LR, 64, device=dev).to(torch.bfloat16) for _ in range(NB)]
    xl = [torch.randint(-128, 128, (bc, LR, LR, 64), device=dev, dtype=torch.int8) for _ in range(NB)]
    t = [torch.empty(bc, LR, LR, 64, device=dev, dtype=torch.bfloat16) for _ in range(NB)]
    pool = torch.empty(bc, LR, 64, device=dev)
    cf, cl = torch.empty(bc, LR, 64, device=dev), torch.empty(bc, LR, 64, device=dev)
    blob = torch.randn(4 * 74 + 4 + 64 * 4 + 64, device=dev) / 8   # QCALayer 'standard' + meta: the bench network's style
    attr = torch.rand(bc, 10, device=dev)
    sq = torch.rand(bc, 64, device=dev) * 0.1
    k = [0]

    def conv1():
        i = k[0] = (k[0] + 1) % NB
        _lib.check(lib.dfir_conv3x3_c64_stats(xh[i].data_ptr(), wp.data_ptr(), bias.data_ptr(), bc, LR, LR, t[i].data_ptr(),
                                              pool.data_ptr(), cf.data_ptr(), cl.data_ptr(), st), "conv1")

    def conv2():
        i = k[0] = (k[0] + 1) % NB
        _lib.check(lib.dfir_conv3x3_c64_scale_skip_hl8(t[i].data_ptr(), wp.data_ptr(), bias.data_ptr(), bc, LR, LR, None,
                                                       xh[i].data_ptr(), xl[i].data_ptr(), xh[i].data_ptr(), xl[i].data_ptr(),
                                                       pool.data_ptr(), cf.data_ptr(), cl.data_ptr(), 1, blob.data_ptr(), 4,
                                                       10, 10, attr.data_ptr(), sq.data_ptr(), 1, st), "conv2")

    def timeit(fn, n=90):
        for _ in range(9):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    for i in range(NB):
        k[0] = i
        conv1()
    ms1 = timeit(conv1)
    ms2 = timeit(conv2)
    flops = bc * LR * LR * CONV64_FLOP_PER_PIXEL
    achieved = flops / (ms1 / 1e3) / 1e12
    conv1_d = {"bound": "tensor", "kernel": "conv3x3_c64_tc_kernel<64, bias+ReLU+statistics> (RCAB conv1, as the schedule launches it)",
               "achieved": round(achieved, 2), "peak": pk["tf_burst"], "unit": "TFLOP/s",
               "frac": round(achieved / pk["tf_burst"], 4), "traffic": NCU["conv1_traffic"],
               "launch_us": round(ms1 * 1e3, 3), "images_per_launch": bc, "in_step_us_ncu": NCU["conv1_in_step_us"],
               "share_of_step_ncu": NCU["conv1_share"], "algorithmic_flops_per_launch": flops,
               "algorithmic_bytes_per_launch": bc * LR * LR * 64 * 4,
               "hbm_gbs": round(bc * LR * LR * 64 * 4 / (ms1 / 1e3) / 1e9, 1),
               "source": NCU["source"], "peak_source": pk["source"] + ", burst (kernel timed alone)"}
    byt = bc * LR * LR * 64 * 8    # t 2 in, hi 2 + lo 1 in, hi 2 + lo 1 out
    gbs = byt / (ms2 / 1e3) / 1e9
    conv2_d = {"bound": "hbm", "kernel": "conv3x3_c64_tc_kernel<64, scale+skip on the hi / 8-bit lo stream> (RCAB conv2 + in-kernel "
                                          "channel/meta attention + scale + residual, as the schedule launches it)",
               "achieved": round(gbs, 1), "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": round(gbs / pk["hbm_gbs"], 4),
               "traffic": NCU["conv2_traffic"], "launch_us": round(ms2 * 1e3, 3), "images_per_launch": bc,
               "in_step_us_ncu": NCU["conv2_in_step_us"], "share_of_step_ncu": NCU["conv2_share"],
               "algorithmic_bytes_per_launch": byt, "tensor_tflops": round(flops / (ms2 / 1e3) / 1e12, 1),
               "source": NCU["source"], "peak_source": pk["source"] + " (STREAM-style copy)"}
    return conv2_d, conv1_d


def other_configs(dev):
    """BASELINE.json configs[2] (Q-EDSR x4, 32 blocks x 256 features, one 480x270 frame) and configs[4] (Q-SAN x4 on
    128x128 images, direct and through the handler's quadrant chop): device time per step, a few steps each, so that the
    driver sees them (N = 1 only)."""
    from SISR.models import ModelInterface
    from deepfir_b200.qrcan import QEDSR
    out = {}

    def timeit(fn, n):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    try:
        torch.manual_seed(8)
        g = torch.Generator().manual_seed(8)
        net = QEDSR(precision="bf16", num_blocks=32, num_features=256, input_para=10, scale=4, res_scale=0.1,
                    q_layer_nonlinearity=False).to(dev).eval()
        x = torch.rand(1, 3, 270, 480, generator=g).to(dev)
        meta = (torch.rand(1, 10, 1, 1, generator=g) * 0.4).to(dev)
        with torch.no_grad():
            ms = timeit(lambda: net(x, meta), 5)
        flop = 100505088 * 270 * 480
        out["qedsr_c3"] = {"workload": "Q-EDSR x4 (32 resblocks, 256 ch) inference, one synthetic 480x270 LR frame",
                           "ms_per_frame": round(ms, 3), "fps": round(1e3 / ms, 2),
                           "output_mpix_s": round(1080 * 1920 / 1e6 / (ms / 1e3), 2),
                           "tflops_algorithmic": round(flop / ms / 1e9, 1),
                           "frac_of_bf16_peak": round(flop / ms / 1e9 / peaks()["tf_burst"], 4)}
        del net
        torch.cuda.empty_cache()
    except Exception as e:  # an auxiliary section must never take the headline line down
        out["qedsr_c3"] = {"error": repr(e)[:200]}
    try:
        torch.manual_seed(8)
        B = 8
        h = ModelInterface.define_model("qsan", device=dev.index or 0, model_save_dir=tempfile.gettempdir(), eval_mode=True,
                                        scale=4, metadata=["blur_kernel"], max_combined_im_size=20000)
        with torch.no_grad():
            for p in (h.net.gamma, h.net.non_local.non_local.W.weight, h.net.non_local.non_local.W.bias):
                p.normal_(0, 0.1)  # the reference zero-initialises these branches; they must do real work here
        g = torch.Generator().manual_seed(8)
        x = torch.rand(B, 3, 128, 128, generator=g)
        meta = torch.rand(B, 10, generator=g, dtype=torch.float64) * 0.4
        keys = [("blur_kernel",) * B] * 10
        xd = x.to(dev)
        attr = h.generate_channels(x, meta, keys).to(dev)
        with torch.no_grad():
            ms_direct = timeit(lambda: h.net(xd, attr), 3)
        xp = x.pin_memory()
        ms_chop = timeit(lambda: h.run_eval(xp, metadata=meta, metadata_keys=keys), 3)
        out["qsan_c5"] = {"workload": "Q-SAN x4 (20 groups x 10 blocks, covariance pooling + Newton-Schulz, region non-local), "
                                      "%d synthetic 128x128 LR images" % B,
                          "ms_direct_forward": round(ms_direct, 3),
                          "output_mpix_s_direct": round(B * 512 * 512 / 1e6 / (ms_direct / 1e3), 2),
                          "ms_handler_run_eval_chop": round(ms_chop, 3),
                          "output_mpix_s_handler": round(B * 512 * 512 / 1e6 / (ms_chop / 1e3), 2),
                          "api": "QSANHandler.run_eval (4 overlapping 74x74 quadrants, host in / host out)"}
        del h
        torch.cuda.empty_cache()
    except Exception as e:
        out["qsan_c5"] = {"error": repr(e)[:200]}
    # training step of the staged networks (Q-SAN 20 x 10, Q-HAN 10 x 20) at configs[3]'s batch shape, through run_train
    for model in ("qsan", "qhan"):
        key = model + "_train"
        try:
            torch.manual_seed(8)
            h = ModelInterface.define_model(model, device=dev.index or 0, model_save_dir=tempfile.gettempdir(), eval_mode=False,
                                            lr=1e-4, scale=4, metadata=["blur_kernel"], precision="bf16")
            g = torch.Generator().manual_seed(8)
            x = torch.rand(16, 3, 64, 64, generator=g).pin_memory()
            y = torch.rand(16, 3, 256, 256, generator=g).pin_memory()
            meta = torch.rand(16, 10, generator=g, dtype=torch.float64) * 0.4
            keys = [("blur_kernel",) * 16] * 10
            losses = []
            step = lambda: losses.append(float(h.run_train(x, y, metadata=meta, metadata_keys=keys)[0]))
            step()
            step()   # (the optimizer re-homes the parameters in its first step: the second forward re-packs once more)
            ms = timeit(step, 3)
            out[key] = {"workload": "%s x4 bf16 training step (published depth), batch 16 x 64x64 LR patches, L1, Adam"
                                    % ("Q-SAN" if model == "qsan" else "Q-HAN"),
                        "ms_per_step": round(ms, 3), "loss_first_last": [round(losses[0], 5), round(losses[-1], 5)],
                        "api": "%sHandler.run_train(x_pinned, y_pinned, metadata=, metadata_keys=)" % model.upper()}
            del h
            torch.cuda.empty_cache()
        except Exception as e:
            out[key] = {"error": repr(e)[:200]}
    return out


def cpu_baseline(sample_images=2, threads=None):
    """The reference's path restated on the CPU (oracle port, fp32 torch-CPU == the ATen kernels the
    reference itself runs on CPU), timed on this box's host cores on a bounded sample."""
    from oracle import deepfir_oracle as O
    from deepfir_b200.qrcan import QRCAN
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    torch.manual_seed(8)
    net = QRCAN(num_metadata=10, **NET_KW)
    sd = {k: v.detach() for k, v in net.state_dict().items()}
    x, meta, _ = synth_batch(sample_images, seed=8)
    attr = meta.float().reshape(sample_images, 10, 1, 1)
    with torch.no_grad():
        O.qrcan_forward(x[:1], attr[:1], sd, style="standard")  # warm-up
        t0 = time.perf_counter()
        O.qrcan_forward(x, attr, sd, style="standard")
        dt = time.perf_counter() - t0
    mpix = sample_images * (LR * SCALE) ** 2 / 1e6
    return {"value": round(mpix / dt, 4), "unit": "MPix/s", "cores": threads, "kind": "port",
            "sample": "%d of the %d images of one step, 1 warm-up image, fp32 torch-CPU oracle (oracle/deepfir_oracle.py)"
                      % (sample_images, IMAGES_PER_GPU), "seconds": round(dt, 3)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    sample = 2
    vals = []
    for _ in range(max(1, min(args.warmup, 1))):
        cpu_baseline(sample_images=1)
    t_all = time.perf_counter()
    for _ in range(max(1, args.steps)):
        vals.append(cpu_baseline(sample_images=sample))
        if time.perf_counter() - t_all > 150:
            break
    v = sum(c["value"] for c in vals) / len(vals)
    cpu = dict(vals[-1])
    cpu["value"] = round(v, 4)
    ms = sample * (LR * SCALE) ** 2 / 1e6 / v * 1e3
    print(json.dumps({
        "impl": "reference", "metric": "Q-RCAN x4 output MPix/s", "value": round(v, 4), "unit": "MPix/s",
        "n_gpus": world, "steps": len(vals), "warmup": 1, "ms_per_step": round(ms, 2), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(workload_config(), precision="fp32 (the reference's arithmetic)", l2="n/a (host CPU)",
                       sample="each step = %d images of the workload on the host CPU" % sample),
        "cpu_baseline": cpu,
        "e2e": {"value": round(v, 4), "unit": "MPix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
