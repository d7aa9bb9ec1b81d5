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
from .sharding import allreduce_mean_


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


class _StepGraphs:
    """CUDA graphs of one training step for a fixed input shape.  A step is ~3000 short kernels: replaying them from
    a graph removes the per-launch host cost and most of the device-side launch gaps.  The forward graph also holds
    the parameter re-pack (`dfir_qrcan_repack`), so it always sees the optimizer's latest weights.  Static buffers:
    x, attributes, output, output gradient; the workspace and the gradient buffer are already persistent."""

    def __init__(self, packed, x, attr):
        B, _, H, W = x.shape
        self.x = torch.empty_like(x)
        self.attr = torch.empty_like(attr)
        self.out = torch.empty(B, packed.out_feats, H * packed.scale, W * packed.scale, device=x.device,
                               dtype=torch.float32)
        self.gout = torch.empty_like(self.out)
        self.ws = packed.train_workspace(B, H, W)
        self.fwd = self.bwd = None
        self.eager_fwd_done = self.eager_bwd_done = False


def _run_forward(lib, packed, x, attr, out, ws, repack):
    B, _, H, W = x.shape
    if repack:
        packed.repack()
    rc = lib.dfir_qrcan_train_forward(C.byref(packed.desc), x.data_ptr(), attr.data_ptr(), out.data_ptr(), B, H, W,
                                      packed.precision, ws.data_ptr(), ws.numel(), _stream(x.device))
    _lib.check(rc, "qrcan_train_forward")


def _run_backward(lib, packed, gstruct, x, attr, gout, ws):
    B, _, H, W = x.shape
    rc = lib.dfir_qrcan_train_backward(C.byref(packed.desc), C.byref(gstruct), x.data_ptr(), attr.data_ptr(),
                                       gout.data_ptr(), B, H, W, packed.precision, ws.data_ptr(), ws.numel(),
                                       _stream(x.device))
    _lib.check(rc, "qrcan_train_backward")


class _QrcanTrain(torch.autograd.Function):
    @staticmethod
    def forward(ctx, anchor, x, attr, net, packed):
        lib = _lib.load_library()
        B, _, H, W = x.shape
        use_graphs = bool(getattr(net, "cuda_graphs", True))
        sg = None
        with torch.cuda.device(x.device):
            if use_graphs:
                key = (B, H, W, x.shape[1], attr.shape[1])
                sg = packed.step_graphs.get(key)
                if sg is None:
                    sg = packed.step_graphs[key] = _StepGraphs(packed, x, attr)
                sg.x.copy_(x)
                sg.attr.copy_(attr)
                if sg.fwd is None and sg.eager_fwd_done:
                    graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(graph):
                        _run_forward(lib, packed, sg.x, sg.attr, sg.out, sg.ws, repack=True)
                    sg.fwd = graph
                if sg.fwd is not None:
                    sg.fwd.replay()
                else:  # first step of this shape runs eagerly (also configures the kernels' attributes)
                    _run_forward(lib, packed, sg.x, sg.attr, sg.out, sg.ws, repack=True)
                    sg.eager_fwd_done = True
                out = sg.out.clone()
                ctx.ws = sg.ws
            else:
                out = torch.empty(B, packed.out_feats, H * packed.scale, W * packed.scale, device=x.device,
                                  dtype=torch.float32)
                ctx.ws = packed.train_workspace(B, H, W)
                _run_forward(lib, packed, x, attr, out, ctx.ws, repack=False)
                ctx.save_for_backward(x, attr)
        ctx.net, ctx.packed, ctx.sg = net, packed, sg
        packed.ws_owner = ctx
        return out

    @staticmethod
    def backward(ctx, gout):
        lib = _lib.load_library()
        packed, net, sg = ctx.packed, ctx.net, ctx.sg
        if packed.ws_owner is not ctx:
            raise RuntimeError("the training workspace was re-used by another forward before this backward ran")
        gout = gout.to(torch.float32).contiguous()
        params = packed.grad_params
        # write into the flat buffer that no live .grad aliases (so accumulation semantics survive when the caller
        # did not zero / drop the gradients between two backward passes)
        first = params[0].grad
        which = 1 if (first is not None and first.data_ptr() == packed.grad_views[0][0].data_ptr()) else 0
        _, gstruct = packed.grad_tables[which]
        dev = gout.device
        with torch.cuda.device(dev):
            if sg is not None and which == 0:
                sg.gout.copy_(gout)
                if sg.bwd is None and sg.eager_bwd_done:
                    graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(graph):
                        _run_backward(lib, packed, gstruct, sg.x, sg.attr, sg.gout, sg.ws)
                    sg.bwd = graph
                if sg.bwd is not None:
                    sg.bwd.replay()
                else:
                    _run_backward(lib, packed, gstruct, sg.x, sg.attr, sg.gout, sg.ws)
                    sg.eager_bwd_done = True
            elif sg is not None:
                _run_backward(lib, packed, gstruct, sg.x, sg.attr, gout, sg.ws)
            else:
                x, attr = ctx.saved_tensors
                _run_backward(lib, packed, gstruct, x, attr, gout, ctx.ws)
        flat = packed.grad_flat[which]
        if getattr(net, "ddp_allreduce", True):
            allreduce_mean_(flat)
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
    if x.requires_grad or attr.requires_grad:
        # the backward pass of the library produces parameter gradients only (the reference's training loop never asks
        # for more, models/__init__.py:466-489): refuse instead of silently returning no gradient for the inputs
        raise RuntimeError("the B200 training path does not compute gradients with respect to the input image or the "
                           "metadata: detach them (x.requires_grad=%s, metadata.requires_grad=%s)"
                           % (x.requires_grad, attr.requires_grad))
    anchor = getattr(net, "_train_anchor", None)
    if anchor is None or anchor.device != x.device:
        anchor = torch.zeros(1, device=x.device, requires_grad=True)
        net._train_anchor = anchor
    return _QrcanTrain.apply(anchor, x, attr, net, packed)
