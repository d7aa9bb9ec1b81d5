"""Helpers for the -m gpu parity tests: everything goes through the C ABI (ctypes), the expected values
come from the CPU oracle."""
import ctypes as C

import numpy as np
import torch

from deepfir_b200 import _lib


def lib():
    return _lib.load_library()


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def sync():
    torch.cuda.synchronize()


def nhwc_bf16(x_nchw):
    return x_nchw.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16).cuda()


def nhwc_f32(x_nchw):
    return x_nchw.permute(0, 2, 3, 1).contiguous().float().cuda()


def to_nchw(x_nhwc):
    return x_nhwc.float().cpu().permute(0, 3, 1, 2).contiguous()


def bf16_round(t):
    return t.to(torch.bfloat16).float()


def pack_bf16(w_oihw, nt_rows=64, co_begin=0, co_stride=1):
    w = w_oihw.float().contiguous().cuda()
    out = torch.empty(9 * nt_rows * 128, dtype=torch.uint8, device="cuda")
    _lib.check(lib().dfir_pack_conv3x3_bf16(w.data_ptr(), out.data_ptr(), w.shape[0], w.shape[1], nt_rows, co_begin,
                                            co_stride, stream()), "pack bf16")
    return out


def pack_f32(w_oihw):
    w = w_oihw.float().contiguous().cuda()
    out = torch.empty(9 * w.shape[1] * w.shape[0], dtype=torch.float32, device="cuda")
    _lib.check(lib().dfir_pack_conv3x3_f32(w.data_ptr(), out.data_ptr(), w.shape[0], w.shape[1], stream()), "pack f32")
    return out


def conv_tc(x_bf16_nhwc, wpacked, bias, epi, cout=64, skip=None, want_f32=False, desc_mode=0, out=None, strides=None,
            out_hw=None):
    """Runs dfir_conv3x3_c64; returns (out_bf16_nhwc | None, out_f32 | None, pool_rows | None)."""
    B, H, W, Ct = x_bf16_nhwc.shape
    bias = bias.float().contiguous().cuda()
    nseg = (W + 127) // 128
    o_bf = o_32 = pool = None
    ps = rs = im = 0
    if epi != 4:
        o_bf = out if out is not None else torch.full((B, H, W, 64), float("nan"), dtype=torch.bfloat16, device="cuda")
        ps, rs, im = strides if strides is not None else (128, W * 128, H * W * 128)
    if epi == 2:
        pool = torch.full((B, nseg, H, 64), float("nan"), dtype=torch.float32, device="cuda")
    if epi == 3 and want_f32:
        o_32 = torch.full((B, H, W, 64), float("nan"), dtype=torch.float32, device="cuda")
    if epi == 4:
        o_32 = torch.full((B, cout, H, W), float("nan"), dtype=torch.float32, device="cuda")
    ptr = lambda t: (t.data_ptr() if t is not None else None)
    rc = lib().dfir_conv3x3_c64(x_bf16_nhwc.data_ptr(), Ct, 0, wpacked.data_ptr(), bias.data_ptr(), B, H, W, epi, cout,
                                ptr(o_bf), ps, rs, im, ptr(skip), ptr(o_32), ptr(pool), desc_mode, stream())
    _lib.check(rc, "conv3x3_c64")
    sync()
    return o_bf, o_32, pool


def max_norm_err(a, b):
    a = a.double()
    b = b.double()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()
