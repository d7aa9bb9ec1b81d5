"""Training step on the B200 path: `QRCAN.forward` under autograd.

The reference trains through `loss.backward()` over the eager modules (`BaseModel.run_train` / `standard_update`,
/root/reference/Code/SISR/models/__init__.py:466-489).  Here the whole network is one autograd node: the forward is
`dfir_qrcan_train_forward` (activations stay in a library workspace), the backward is `dfir_qrcan_train_backward`,
which writes the gradient of EVERY parameter into one flat fp32 buffer; each parameter's `.grad` is a view of it.
The optimizer (torch Adam), the L1 criterion and the schedulers are the reference's own objects and see ordinary
`.grad` tensors.  Under torch.distributed (one process per GPU) the flat buffer is all-reduced once per step over
NCCL — the only exchange step of data-parallel training.
"""
import ctypes as C

import torch

from . import _lib


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


class _QrcanTrain(torch.autograd.Function):
    @staticmethod
    def forward(ctx, anchor, x, attr, net, packed):
        lib = _lib.load_library()
        B, _, H, W = x.shape
        out = torch.empty(B, packed.out_feats, H * packed.scale, W * packed.scale, device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device):
            ws = packed.train_workspace(B, H, W)
            rc = lib.dfir_qrcan_train_forward(C.byref(packed.desc), x.data_ptr(), attr.data_ptr(), out.data_ptr(), B, H,
                                              W, packed.precision, ws.data_ptr(), ws.numel(), _stream(x.device))
        _lib.check(rc, "qrcan_train_forward")
        ctx.net, ctx.packed, ctx.ws = net, packed, ws
        ctx.save_for_backward(x, attr)
        return out

    @staticmethod
    def backward(ctx, gout):
        lib = _lib.load_library()
        packed, net = ctx.packed, ctx.net
        x, attr = ctx.saved_tensors
        B, _, H, W = x.shape
        gout = gout.to(torch.float32).contiguous()
        params = packed.grad_params
        # write into the flat buffer that no live .grad aliases (so accumulation semantics survive when the caller
        # did not zero / drop the gradients between two backward passes)
        first = params[0].grad
        which = 1 if (first is not None and first.data_ptr() == packed.grad_views[0][0].data_ptr()) else 0
        _, gstruct = packed.grad_tables[which]
        with torch.cuda.device(x.device):
            if packed.train_workspace(B, H, W) is not ctx.ws:
                raise RuntimeError("the training workspace was re-used by another forward before backward ran")
            rc = lib.dfir_qrcan_train_backward(C.byref(packed.desc), C.byref(gstruct), x.data_ptr(), attr.data_ptr(),
                                               gout.data_ptr(), B, H, W, packed.precision, ctx.ws.data_ptr(),
                                               ctx.ws.numel(), _stream(x.device))
        _lib.check(rc, "qrcan_train_backward")
        flat = packed.grad_flat[which]
        if getattr(net, "ddp_allreduce", True) and torch.distributed.is_available() \
                and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1:
            torch.distributed.all_reduce(flat)
            flat.div_(torch.distributed.get_world_size())
        views = packed.grad_views[which]
        acc_p, acc_g = [], []
        for p, g in zip(params, views):
            if not p.requires_grad:
                continue
            if p.grad is None:
                p.grad = g
            else:
                acc_p.append(p.grad)
                acc_g.append(g)
        if acc_p:
            torch._foreach_add_(acc_p, acc_g)
        return gout.new_zeros(1), None, None, None, None


def qrcan_train_apply(net, packed, x, attr):
    anchor = getattr(net, "_train_anchor", None)
    if anchor is None or anchor.device != x.device:
        anchor = torch.zeros(1, device=x.device, requires_grad=True)
        net._train_anchor = anchor
    return _QrcanTrain.apply(anchor, x, attr, net, packed)
