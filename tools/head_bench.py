import ctypes as C, sys, os
sys.path.insert(0, "super-resolution-meta-attention-networks_b200")
import torch
from deepfir_b200 import _lib
lib = _lib.load_library()
x = torch.rand(32, 3, 128, 128, device="cuda")
w = torch.randn(27 * 64, device="cuda"); b = torch.zeros(64, device="cuda")
o32 = torch.empty(32, 128, 128, 64, device="cuda"); obf = torch.empty(32, 128, 128, 64, device="cuda", dtype=torch.bfloat16)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
def run():
    _lib.check(lib.dfir_head_conv(x.data_ptr(), w.data_ptr(), b.data_ptr(), o32.data_ptr(), obf.data_ptr(), 32, 3, 128, 128, 64, st), "head")
for _ in range(5): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50): run()
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) / 50 * 1e3
print("head conv 32x128x128: %.1f us, %.1f GB/s written (6 B/elem)" % (us, 32 * 128 * 128 * 64 * 6 / us / 1e3))
