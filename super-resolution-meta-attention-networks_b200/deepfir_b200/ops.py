"""torch custom ops over the C ABI (namespace ``dfir``).  CUDA-only: no CPU kernels are registered, so a
CPU tensor fails loudly inside the dispatcher instead of silently falling back."""
import ctypes as C

import torch

from . import _lib
from .qrcan import packed_from_handle


def _stream(t):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


@torch.library.custom_op("dfir::qrcan_forward", mutates_args=(), device_types="cuda")
def qrcan_forward(x: torch.Tensor, attributes: torch.Tensor, handle: int, precision: int) -> torch.Tensor:
    """QRCAN.forward(x, metadata) (reference attention_manipulators/architectures.py:309-316)."""
    lib = _lib.load_library()
    p = packed_from_handle(handle)
    B, _, H, W = x.shape
    out = torch.empty(B, p.out_feats, H * p.scale, W * p.scale, device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        ws = p.workspace(B, H, W, precision)
        rc = lib.dfir_qrcan_forward(C.byref(p.desc), x.data_ptr(), attributes.data_ptr(), out.data_ptr(), B, H, W,
                                    precision, ws.data_ptr(), ws.numel(), _stream(x))
    _lib.check(rc, "qrcan_forward")
    return out


@qrcan_forward.register_fake
def _(x, attributes, handle, precision):
    p = packed_from_handle(handle)
    B, _, H, W = x.shape
    return x.new_empty(B, p.out_feats, H * p.scale, W * p.scale, dtype=torch.float32)
