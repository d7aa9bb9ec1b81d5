"""Pipeline trace of one CTA of the conv kernel (needs `make PROBES=1`): python tools/trace_conv.py [conv1|stats|ss]
Prints, per row, when each role reached its events (SM clocks relative to the first event) and the row period."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "super-resolution-meta-attention-networks_b200"))
os.environ["DFIR_DEBUG_PROBE"] = str(32768 | int(os.environ.get("EXTRA_PROBE", "0")))
import torch
from deepfir_b200 import _lib

lib = _lib.load_library()
dev = torch.device("cuda")
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
mode = (sys.argv[1:] or ["conv1"])[0]
bc, LR = int(os.environ.get("BC", "32")), 128
wp = torch.empty(9 * 64 * 128, dtype=torch.uint8, device=dev)
w = (torch.randn(64, 64, 3, 3, device=dev) / 24).contiguous()
bias = torch.zeros(64, device=dev)
_lib.check(lib.dfir_pack_conv3x3_bf16(w.data_ptr(), wp.data_ptr(), 64, 64, 64, 0, 1, st()), "pack")
a = torch.randn(bc, LR, LR, 64, device=dev).to(torch.bfloat16)
b = torch.empty_like(a)
x = torch.randn(bc, LR, LR, 64, device=dev)
pool = torch.empty(bc, LR, 64, device=dev); cf = torch.empty(bc, LR, 64, device=dev); cl = torch.empty(bc, LR, 64, device=dev)
sv = torch.rand(bc, 64, device=dev)
xh = torch.randn(bc, LR, LR, 64, device=dev).to(torch.bfloat16)
xl = (torch.randn(bc, LR, LR, 64, device=dev) * 1e-3).to(torch.bfloat16)
xl8 = torch.randint(-128, 128, (bc, LR, LR, 64), device=dev, dtype=torch.int8)
ist = torch.zeros(bc, 9, 64, device=dev, dtype=torch.int64)
pool.normal_(); cf.normal_(); cl.normal_()
blob = torch.randn(4 * 74 + 4 + 64 * 4 + 64, device=dev) / 8
attr = torch.rand(bc, 10, device=dev); sq = torch.rand(bc, 64, device=dev) * 0.1
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def launch():
    if mode == "conv1":
        _lib.check(lib.dfir_conv3x3_c64(a.data_ptr(), 64, 0, wp.data_ptr(), bias.data_ptr(), bc, LR, LR, 1, 64, b.data_ptr(),
                                        128, LR * 128, LR * LR * 128, None, None, pool.data_ptr(), 0, st()), "conv")
    elif mode == "stats":
        _lib.check(lib.dfir_conv3x3_c64_stats(a.data_ptr(), wp.data_ptr(), bias.data_ptr(), bc, LR, LR, b.data_ptr(),
                                              pool.data_ptr(), cf.data_ptr(), cl.data_ptr(), st()), "c1")
    elif mode in ("statsfx", "statsw"):
        ist.zero_()
        _lib.check(lib.dfir_conv3x3_c64_stats_fx(a.data_ptr(), wp.data_ptr(), bias.data_ptr(), bc, LR, LR, b.data_ptr(),
                                                 ist.data_ptr(), 1 if mode == "statsw" else 0, st()), "c1fx")
    elif mode == "sshl8fx":   # statistics as the fixed-point image sums (what the network schedule launches)
        _lib.check(lib.dfir_conv3x3_c64_scale_skip_hl8(a.data_ptr(), wp.data_ptr(), bias.data_ptr(), bc, LR, LR,
                                                       None, xh.data_ptr(), xl8.data_ptr(), xh.data_ptr(), xl8.data_ptr(),
                                                       ist.data_ptr(), None, None, 1, blob.data_ptr(), 4, 10, 10,
                                                       attr.data_ptr(), sq.data_ptr(), int(os.environ.get("DESC", "0")), st()), "sshl8fx")
    elif mode in ("sshl8", "sshl8stats"):
        stats = mode == "sshl8stats"
        _lib.check(lib.dfir_conv3x3_c64_scale_skip_hl8(a.data_ptr(), wp.data_ptr(), bias.data_ptr(), bc, LR, LR,
                                                       None if stats else sv.data_ptr(), xh.data_ptr(), xl8.data_ptr(),
                                                       xh.data_ptr(), xl8.data_ptr(), pool.data_ptr() if stats else None,
                                                       cf.data_ptr() if stats else None, cl.data_ptr() if stats else None,
                                                       1 if stats else 0, blob.data_ptr() if stats else None, 4, 10, 10,
                                                       attr.data_ptr() if stats else None, sq.data_ptr() if stats else None,
                                                       int(os.environ.get("DESC", "0")), st()), "sshl8")
    elif mode in ("sshl", "sshlstats"):
        stats = mode == "sshlstats"
        _lib.check(lib.dfir_conv3x3_c64_scale_skip_hl(a.data_ptr(), wp.data_ptr(), bias.data_ptr(), bc, LR, LR,
                                                      None if stats else sv.data_ptr(), xh.data_ptr(), xl.data_ptr(),
                                                      xh.data_ptr(), xl.data_ptr(), pool.data_ptr() if stats else None,
                                                      cf.data_ptr() if stats else None, cl.data_ptr() if stats else None,
                                                      1 if stats else 0, blob.data_ptr() if stats else None, 4, 10, 10,
                                                      attr.data_ptr() if stats else None, sq.data_ptr() if stats else None,
                                                      int(os.environ.get("DESC", "0")), st()), "sshl")
    else:
        _lib.check(lib.dfir_conv3x3_c64_scale_skip(a.data_ptr(), wp.data_ptr(), bias.data_ptr(), bc, LR, LR, sv.data_ptr(),
                                                   x.data_ptr(), x.data_ptr(), b.data_ptr(), None, None, None, 0, None, 4,
                                                   10, 10, None, None, st()), "ss")


for _ in range(3):
    launch()
if os.environ.get("FLUSH", "1") == "1":
    flush.zero_()
torch.cuda.synchronize()
launch()
buf = (C.c_ulonglong * 1024)()
_lib.check(lib.dfir_debug_trace(C.byref(buf)), "trace")
tr = [[buf[k * 64 + i] for i in range(64)] for k in range(16)]
t0 = min(v for row in tr for v in row if v)
names = ["tma_issue", "ld_full", "ld_aempty", "ld_done", "mma_top", "mma_24", "mma_waited", "mma_36", "e0_wait", "e0_tfull",
         "e0_release", "e1_wait", "e1_tfull", "e1_release", "e0_end", "e1_end"]
if int(os.environ.get("EXTRA_PROBE", "0")) & 262144:   # tile loop of epilogue group 0 (first tile of each row)
    names[11:14] = ["t_issued", "t_landed", "t_updated"]
    names[15] = "t2_fenced"
if mode in ("conv1", "stats", "statsfx"):
    names[11:14] = ["e0_stfree", "e0_staged", "e0_sums"]
    names[15] = "e0_bar2"
print("mode %s bc %d (SM clocks relative to the first event; 0 = not recorded)" % (mode, bc))
print("row " + " ".join("%10s" % n for n in names))
for i in range(34):
    print("%3d " % i + " ".join("%10d" % (tr[k][i] - t0 if tr[k][i] else 0) for k in range(16)))
pro = [tr[8][i] for i in range(40, 50)]
if any(pro):
    print("prologue of epilogue group 0 (clk after griddepcontrol.wait): " + " ".join(
        "%s=%d" % (n, v - pro[0]) for n, v in zip(["cap_bar", "weights", "stats_loaded", "S_stored", "matvec", "y", "attn", "img_done",
                                                   "both_groups"], pro[1:]) if v) + "  first_row_wait=%d" % (tr[8][0] - pro[0]))
if int(os.environ.get("EXTRA_PROBE", "0")) & 262144 and any(tr[2][i] for i in range(32, 46)):
    # stamps inside the update of the first pixel slot of a row (epilogue group 0, warp 0): kinds 0 / 2 / 3 / 5 at rows 32..,
    # kind 1 = loads issued of the LAST pixel slot
    print("update of pixel slot 0 (clk after the tile landed): loads issued, loads landed, arithmetic done, stores issued | "
          "slot 3 loads issued | half 0 done, row fenced")
    for r in range(14):
        if tr[2][32 + r] and tr[12][r]:
            base = tr[12][r]
            print("  urow %2d: %5d %5d %5d %5d | %5d | %5d %5d" % (r, tr[0][32 + r] - base, tr[2][32 + r] - base, tr[3][32 + r] - base,
                                                               tr[5][32 + r] - base, tr[1][32 + r] - base, tr[13][r] - base,
                                                               tr[15][r] - base))
for k in (0, 3, 4, 7):
    v = [tr[k][i] for i in range(4, 27) if tr[k][i]]
    if len(v) > 2:
        print("%-10s mean period over rows 4..26: %.0f clk" % (names[k], (v[-1] - v[0]) / (len(v) - 1)))
