"""Micro-benchmarks on the GPU box: trunk conv and scale-residual vs images per launch, and the full
forward vs chunk size.  Prints one line per measurement."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "super-resolution-meta-attention-networks_b200"))
import torch
from deepfir_b200 import _lib
from deepfir_b200.qrcan import QRCAN

lib = _lib.load_library()
dev = torch.device("cuda")
LR = int(os.environ.get("LR", "128"))
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)


def timeit(fn, n=60, warm=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3  # us


BCS = tuple(int(t) for t in os.environ.get("BC", "1,2,4,7,8,9,16,32,64").split(","))
NREP = [60, 10]


def conv_sweep(epi):
    wp = torch.empty(9 * 64 * 128, dtype=torch.uint8, device=dev)
    w = (torch.randn(64, 64, 3, 3, device=dev) / 24).contiguous()
    bias = torch.zeros(64, device=dev)
    _lib.check(lib.dfir_pack_conv3x3_bf16(w.data_ptr(), wp.data_ptr(), 64, 64, 64, 0, 1, st()), "pack")
    for bc in BCS:
        a = torch.randn(bc, LR, LR, 64, device=dev).to(torch.bfloat16)
        b = torch.empty_like(a)
        pool = torch.empty(bc, LR, 64, device=dev)
        flip = [a, b]

        def launch():
            src, dst = flip
            _lib.check(lib.dfir_conv3x3_c64(src.data_ptr(), 64, 0, wp.data_ptr(), bias.data_ptr(), bc, LR, LR, epi, 64,
                                            dst.data_ptr(), 128, LR * 128, LR * LR * 128, None, None, pool.data_ptr(),
                                            0, st()), "conv")
            flip.reverse()
        us = timeit(launch, NREP[0], NREP[1])
        fl = bc * LR * LR * 2 * 64 * 64 * 9
        rows = bc * LR / 148.0
        print("conv epi=%d bc=%3d: %8.2f us  %7.1f TFLOP/s  rows/CTA %.2f  us/row-slot %.3f" %
              (epi, bc, us, fl / us / 1e6, rows, us / max(1.0, -(-bc * LR // 148))))


def fused_sweep():
    wp = torch.empty(9 * 64 * 128, dtype=torch.uint8, device=dev)
    w = (torch.randn(64, 64, 3, 3, device=dev) / 24).contiguous()
    bias = torch.zeros(64, device=dev)
    _lib.check(lib.dfir_pack_conv3x3_bf16(w.data_ptr(), wp.data_ptr(), 64, 64, 64, 0, 1, st()), "pack")
    blob = torch.randn(4 * 64 + 4 + 64 * 4 + 64, device=dev) / 8
    for bc in BCS:
        r = torch.randn(bc, LR, LR, 64, device=dev).to(torch.bfloat16)
        x = torch.randn(bc, LR, LR, 64, device=dev)
        xo = torch.empty_like(x)
        t = torch.empty_like(r)
        pool = torch.randn(bc, LR, 64, device=dev)
        attr = torch.rand(bc, 10, device=dev)
        sq = torch.rand(bc, 64, device=dev)

        def launch():
            _lib.check(lib.dfir_conv3x3_c64_fused(r.data_ptr(), x.data_ptr(), xo.data_ptr(), pool.data_ptr(), 1,
                                                  blob.data_ptr(), 4, 10, 10, attr.data_ptr(), sq.data_ptr(), 1.0,
                                                  wp.data_ptr(), bias.data_ptr(), bc, LR, LR, 1, t.data_ptr(), None,
                                                  None, st()), "fused")
        us = timeit(launch, NREP[0], NREP[1])
        fl = bc * LR * LR * 2 * 64 * 64 * 9
        byt = bc * LR * LR * 64 * 12
        print("fused conv bc=%3d: %8.2f us  %7.1f TFLOP/s  %7.1f GB/s (12 B/elem)" % (bc, us, fl / us / 1e6, byt / us / 1e3))


def ss_sweep():
    wp = torch.empty(9 * 64 * 128, dtype=torch.uint8, device=dev)
    w = (torch.randn(64, 64, 3, 3, device=dev) / 24).contiguous()
    bias = torch.zeros(64, device=dev)
    _lib.check(lib.dfir_pack_conv3x3_bf16(w.data_ptr(), wp.data_ptr(), 64, 64, 64, 0, 1, st()), "pack")
    for bc in BCS:
        t = torch.randn(bc, LR, LR, 64, device=dev).to(torch.bfloat16)
        x = torch.randn(bc, LR, LR, 64, device=dev)
        xb = torch.empty_like(t)
        sv = torch.rand(bc, 64, device=dev)

        def launch():
            _lib.check(lib.dfir_conv3x3_c64_scale_skip(t.data_ptr(), wp.data_ptr(), bias.data_ptr(), bc, LR, LR,
                                                       sv.data_ptr(), x.data_ptr(), x.data_ptr(), xb.data_ptr(), None,
                                                       None, None, 0, None, 4, 10, 10, None, None, st()), "ss")
        us = timeit(launch, NREP[0], NREP[1])
        byt = bc * LR * LR * 64 * 12
        print("conv+scale_skip bc=%3d: %8.2f us  %7.1f GB/s (12 B/elem)" % (bc, us, byt / us / 1e3))


def ss_hl_sweep():
    wp = torch.empty(9 * 64 * 128, dtype=torch.uint8, device=dev)
    w = (torch.randn(64, 64, 3, 3, device=dev) / 24).contiguous()
    bias = torch.zeros(64, device=dev)
    _lib.check(lib.dfir_pack_conv3x3_bf16(w.data_ptr(), wp.data_ptr(), 64, 64, 64, 0, 1, st()), "pack")
    for bc in BCS:
        t = torch.randn(bc, LR, LR, 64, device=dev).to(torch.bfloat16)
        xh = torch.randn(bc, LR, LR, 64, device=dev).to(torch.bfloat16)
        xl = (torch.randn(bc, LR, LR, 64, device=dev) * 1e-3).to(torch.bfloat16)
        sv = torch.rand(bc, 64, device=dev) * 0.1

        def launch():
            _lib.check(lib.dfir_conv3x3_c64_scale_skip_hl(t.data_ptr(), wp.data_ptr(), bias.data_ptr(), bc, LR, LR,
                                                          sv.data_ptr(), xh.data_ptr(), xl.data_ptr(), xh.data_ptr(),
                                                          xl.data_ptr(), None, None, None, 0, None, 4, 10, 10, None, None,
                                                          int(os.environ.get("DESC", "0")), st()), "ss hl")
        us = timeit(launch, NREP[0], NREP[1])
        byt = bc * LR * LR * 64 * 10
        print("conv+scale_skip hi/lo stream bc=%3d: %8.2f us  %7.1f GB/s (10 B/elem)" % (bc, us, byt / us / 1e3))


def sr_sweep():
    for bc in (1, 4, 7, 8, 16, 32):
        r = torch.randn(bc, LR, LR, 64, device=dev).to(torch.bfloat16)
        x = torch.randn(bc, LR, LR, 64, device=dev)
        xb = torch.empty(bc, LR, LR, 64, device=dev, dtype=torch.bfloat16)
        pool = torch.randn(bc, LR, 64, device=dev)
        blob = torch.randn(4 * 64 + 4 + 64 * 4 + 64, device=dev) / 8
        attr = torch.rand(bc, 10, device=dev)
        sq = torch.rand(bc, 64, device=dev)

        def launch():
            _lib.check(lib.dfir_ca_scale_residual(r.data_ptr(), 1, x.data_ptr(), pool.data_ptr(), LR, 1, blob.data_ptr(),
                                                  64, 4, 10, 10, attr.data_ptr(), sq.data_ptr(), 1.0, x.data_ptr(),
                                                  xb.data_ptr(), bc, LR, LR, st()), "sr")
        us = timeit(launch)
        byt = bc * LR * LR * 64 * 12
        print("scale_residual bc=%3d: %8.2f us  %7.1f GB/s (12 B/elem algorithmic)" % (bc, us, byt / us / 1e3))


def forward_sweep():
    torch.manual_seed(8)
    B = 32
    x = torch.rand(B, 3, LR, LR, device=dev)
    meta = torch.rand(B, 10, 1, 1, device=dev) * 0.4
    for chunk in tuple(int(t) for t in os.environ.get("CHUNKS", "8,12,16,24,32").split(",")):
        net = QRCAN(n_resgroups=10, n_resblocks=20, style="standard", num_metadata=10, include_q_layer=True,
                    precision="bf16", chunk_images=chunk).to(dev).eval()
        with torch.no_grad():
            us = timeit(lambda: net(x, meta), n=5, warm=2)
        print("forward B=32 chunk=%2d: %8.2f ms  %7.1f MPix/s" % (chunk, us / 1e3, B * (4 * LR) ** 2 / us))
        del net


which = sys.argv[1:] or ["conv", "sr", "fwd"]
if "conv" in which:
    conv_sweep(1)
    conv_sweep(2)
if "ss" in which:
    ss_sweep()
if "sshl" in which:
    ss_hl_sweep()
if "fused" in which:
    fused_sweep()
if "sr" in which:
    sr_sweep()
if "fwd" in which:
    forward_sweep()
if "one" in which:  # few launches of selected shapes, for ncu (BC=32 python tools/conv_bench.py one)
    NREP[:] = [4, 2]
    conv_sweep(int(os.environ.get("EPI", "1")))


def wgrad_bench():
    B, H, W = (int(t) for t in os.environ.get("WG", "16,64,64").split(","))
    x = torch.randn(B, H, W, 64, device=dev).to(torch.bfloat16)
    dy = torch.randn(B, H, W, 64, device=dev).to(torch.bfloat16)
    n = lib.dfir_conv3x3_wgrad_scratch_bytes(B, H, W, 64, 64, 0)
    scratch = torch.empty(n, dtype=torch.uint8, device=dev)
    dw = torch.empty(64, 64, 3, 3, device=dev); db = torch.empty(64, device=dev)

    def launch():
        _lib.check(lib.dfir_conv3x3_wgrad_c64(dy.data_ptr(), 0, 0, 0, x.data_ptr(), B, H, W, dw.data_ptr(), db.data_ptr(), 0, 1,
                                              scratch.data_ptr(), n, st()), "wgrad")
    us = timeit(launch, 100, 10)
    fl = B * H * W * 2 * 64 * 64 * 9
    print("wgrad(+reduce unless probe&8) %dx%dx%d probe=%s: %7.2f us  %6.1f TFLOP/s" %
          (B, H, W, os.environ.get("DFIR_WGRAD_PROBE", "0"), us, fl / us / 1e6))


if "wgrad" in which:
    wgrad_bench()


def skip_direct_bench():
    """EPI_BIAS_SKIP (epi 3): lane-major epilogue without the shared-memory transposition tile"""
    wp = torch.empty(9 * 64 * 128, dtype=torch.uint8, device=dev)
    w = (torch.randn(64, 64, 3, 3, device=dev) / 24).contiguous()
    bias = torch.zeros(64, device=dev)
    _lib.check(lib.dfir_pack_conv3x3_bf16(w.data_ptr(), wp.data_ptr(), 64, 64, 64, 0, 1, st()), "pack")
    for bc in BCS:
        t = torch.randn(bc, LR, LR, 64, device=dev).to(torch.bfloat16)
        x = torch.randn(bc, LR, LR, 64, device=dev)
        xb = torch.empty_like(t)

        def launch():
            _lib.check(lib.dfir_conv3x3_c64(t.data_ptr(), 64, 0, wp.data_ptr(), bias.data_ptr(), bc, LR, LR, 3, 64,
                                            xb.data_ptr(), 128, LR * 128, LR * LR * 128, x.data_ptr(), x.data_ptr(), None,
                                            0, st()), "conv epi3")
        us = timeit(launch, NREP[0], NREP[1])
        print("conv+bias_skip (lane-major, no tile) bc=%3d: %8.2f us   %7.1f GB/s (12 B/elem)" % (bc, us, bc * LR * LR * 64 * 12 / us / 1e3))


if "skipdirect" in which:
    skip_direct_bench()


def ss_stats_bench():
    """RCAB conv2 as the default schedule launches it: attention vector evaluated in the kernel from conv1's statistics"""
    wp = torch.empty(9 * 64 * 128, dtype=torch.uint8, device=dev)
    w = (torch.randn(64, 64, 3, 3, device=dev) / 24).contiguous()
    bias = torch.zeros(64, device=dev)
    _lib.check(lib.dfir_pack_conv3x3_bf16(w.data_ptr(), wp.data_ptr(), 64, 64, 64, 0, 1, st()), "pack")
    blob = torch.randn(4 * 64 + 4 + 64 * 4 + 64, device=dev) / 8
    for bc in BCS:
        t = torch.randn(bc, LR, LR, 64, device=dev).to(torch.bfloat16)
        x = torch.randn(bc, LR, LR, 64, device=dev)
        xb = torch.empty_like(t)
        pool = torch.randn(bc, LR, 64, device=dev); cf = torch.randn(bc, LR, 64, device=dev); cl = torch.randn(bc, LR, 64, device=dev)
        attr = torch.rand(bc, 10, device=dev); sq = torch.rand(bc, 64, device=dev)

        def launch():
            _lib.check(lib.dfir_conv3x3_c64_scale_skip(t.data_ptr(), wp.data_ptr(), bias.data_ptr(), bc, LR, LR, None,
                                                       x.data_ptr(), x.data_ptr(), xb.data_ptr(), pool.data_ptr(),
                                                       cf.data_ptr(), cl.data_ptr(), 1, blob.data_ptr(), 4, 10, 10,
                                                       attr.data_ptr(), sq.data_ptr(), st()), "ss stats")
        us = timeit(launch, NREP[0], NREP[1])
        print("conv+scale_skip WITH in-kernel attention bc=%3d: %8.2f us" % (bc, us))


if "ssstats" in which:
    ss_stats_bench()


def pair_bench():
    """One RCAB exactly as the default schedule chains it: conv1 (+ReLU+statistics) xb -> t, then conv2 (+in-kernel
    attention + scale + skip) t, x -> x, xb — versus each kernel alone on buffers rotated to defeat the L2."""
    wp = torch.empty(9 * 64 * 128, dtype=torch.uint8, device=dev)
    w = (torch.randn(64, 64, 3, 3, device=dev) / 48).contiguous()
    bias = torch.zeros(64, device=dev)
    _lib.check(lib.dfir_pack_conv3x3_bf16(w.data_ptr(), wp.data_ptr(), 64, 64, 64, 0, 1, st()), "pack")
    blob = torch.randn(4 * 74 + 4 + 64 * 4 + 64, device=dev) / 8
    for bc in BCS:
        NB = 4
        xs = [torch.randn(bc, LR, LR, 64, device=dev) for _ in range(NB)]
        xbs = [x.to(torch.bfloat16) for x in xs]
        ts = [torch.empty_like(xbs[0]) for _ in range(NB)]
        pool = torch.empty(bc, LR, 64, device=dev); cf = torch.empty(bc, LR, 64, device=dev); cl = torch.empty(bc, LR, 64, device=dev)
        attr = torch.rand(bc, 10, device=dev); sq = torch.rand(bc, 64, device=dev) * 0.1
        k = [0]

        def conv1(i):
            _lib.check(lib.dfir_conv3x3_c64_stats(xbs[i].data_ptr(), wp.data_ptr(), bias.data_ptr(), bc, LR, LR,
                                                  ts[i].data_ptr(), pool.data_ptr(), cf.data_ptr(), cl.data_ptr(), st()), "c1")

        def conv2(i):
            _lib.check(lib.dfir_conv3x3_c64_scale_skip(ts[i].data_ptr(), wp.data_ptr(), bias.data_ptr(), bc, LR, LR, None,
                                                       xs[i].data_ptr(), xs[i].data_ptr(), xbs[i].data_ptr(), pool.data_ptr(),
                                                       cf.data_ptr(), cl.data_ptr(), 3, blob.data_ptr(), 4, 10, 10,
                                                       attr.data_ptr(), sq.data_ptr(), st()), "c2")

        def pair_same():
            conv1(0); conv2(0)

        def c1_rot():
            k[0] = (k[0] + 1) % NB; conv1(k[0])

        def c2_rot():
            k[0] = (k[0] + 1) % NB; conv2(k[0])

        conv1(0)
        for name, fn in (("conv1 same buffers", lambda: conv1(0)), ("conv1 rotating 4 buffers", c1_rot),
                         ("conv2 same buffers", lambda: conv2(0)), ("conv2 rotating 4 buffers", c2_rot),
                         ("pair conv1->conv2 (the chain)", pair_same)):
            us = timeit(fn, NREP[0], NREP[1])
            print("bc=%3d policy=%s pdl=%s  %-32s %8.2f us" % (bc, os.environ.get("DFIR_L2_POLICY", "default"),
                                                             os.environ.get("DFIR_PDL", "default"), name, us), flush=True)


if "pair" in which:
    pair_bench()
