"""Training step of Q-SAN / Q-HAN (and the non-meta SAN / HAN) on the B200 path.

The reference trains these networks through autograd over their eager forward (`BaseModel.run_train`,
/root/reference/Code/SISR/models/__init__.py:466-489 over attention_manipulators/architectures.py:447-467 and :514-540).
Here the whole network is ONE autograd node, as for Q-RCAN (deepfir_b200/train.py): its forward runs the library's staged
training entry points for the shared conv trunk (`dfir_qrcan_train_stage_forward`: head, residual groups with the
activation stash, upsampler + tail) with the networks' own layers in between, its backward runs the matching
`dfir_qrcan_train_stage_backward` calls and the backward operators of those layers (non-local attention, covariance
pooling + Newton-Schulz square root + SOCA MLP, LAM, CSAM, the group / fusion convs).  Every parameter's gradient lands in
the flat gradient buffer of `PackedQrcan.enable_training`; `.grad` tensors are views of it and the optimizer, criterion and
schedulers stay the reference's own torch objects.  Feature maps between stages are fp32 NHWC device tensors.

The trunk convs follow the network's `precision`; so do the 3x3 convs outside the trunk (bf16 mode: tensor-core operators per
64-channel input chunk, fp32 mode: CUDA-core kernels).  The attention layers outside the trunk are fp32 in both modes.
"""
import ctypes as C

import torch

from . import _lib
from .sharding import allreduce_mean_

T_HEAD, T_GROUPS, T_TAIL, T_ATTN = 1, 2, 8, 16


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _ptr(t):
    return None if t is None else t.data_ptr()


class _Ops:
    """thin wrappers of the C ABI calls a staged step makes (one network, one device, one step)"""

    def __init__(self, net, packed, B, H, W):
        self.lib = _lib.load_library()
        self.net, self.pk = net, packed
        self.B, self.H, self.W, self.C = B, H, W, packed.desc.n_feats
        self.dev = packed.device
        self.st = _stream(self.dev)
        self.ws = packed.train_workspace(B, H, W)
        self.f32 = dict(device=self.dev, dtype=torch.float32)
        self._scratch = {}

    # ---- buffers
    def feat(self, C_=None):
        return torch.empty(self.B, self.H, self.W, C_ or self.C, **self.f32)

    def scratch(self, key, nbytes):
        t = self._scratch.get(key)
        if t is None or t.numel() < nbytes:
            t = self._scratch[key] = torch.empty(max(int(nbytes), 16), device=self.dev, dtype=torch.uint8)
        return t

    # ---- trunk stages
    def stage_fwd(self, stage, g0=0, g1=0, x=None, attr=None, feat_in=None, feat_out=None, group_out=None, out=None):
        pk = self.pk
        rc = self.lib.dfir_qrcan_train_stage_forward(C.byref(pk.desc), stage, g0, g1, _ptr(x), _ptr(attr), _ptr(feat_in),
                                                     _ptr(feat_out), _ptr(group_out), _ptr(out), self.B, self.H, self.W,
                                                     pk.precision, self.ws.data_ptr(), self.ws.numel(), self.st)
        _lib.check(rc, "qrcan_train_stage_forward(%d)" % stage)

    def stage_bwd(self, gstruct, stage, g0=0, g1=0, x=None, attr=None, gout=None, gfeat_out=None, gfeat_in=None):
        pk = self.pk
        rc = self.lib.dfir_qrcan_train_stage_backward(C.byref(pk.desc), C.byref(gstruct), stage, g0, g1, _ptr(x), _ptr(attr),
                                                      _ptr(gout), _ptr(gfeat_out), _ptr(gfeat_in), self.B, self.H, self.W,
                                                      pk.precision, self.ws.data_ptr(), self.ws.numel(), self.st)
        _lib.check(rc, "qrcan_train_stage_backward(%d)" % stage)

    # ---- 3x3 convs outside the trunk.  bf16 mode, 64 output channels, Cin a multiple of 64: the tensor-core operators, one
    # launch per 64-channel input chunk (forward: fp32 running sum chained through the skip input; data gradient: one
    # transposed-weight conv per chunk; weight gradient: one tcgen05 wgrad per chunk).  Otherwise the fp32 CUDA-core kernels.
    def _tc(self, conv):
        return self.pk.precision == 0 and conv.out_channels == 64 and conv.in_channels % 64 == 0

    def _chunks_bf16(self, x):
        nc = x.shape[-1] // 64
        xb = x.to(torch.bfloat16)
        return [xb] if nc == 1 else [xb[..., i * 64:(i + 1) * 64].contiguous() for i in range(nc)]

    def conv_fwd(self, conv, x, skip=None):
        cin, cout = conv.in_channels, conv.out_channels
        if self._tc(conv):
            nc = cin // 64
            w = conv.weight.detach()
            out = self.feat(64)
            junk = torch.empty(self.B, self.H, self.W, 64, device=self.dev, dtype=torch.bfloat16)
            for i, chunk in enumerate(self._chunks_bf16(x)):
                wi = w[:, i * 64:(i + 1) * 64].contiguous()
                t = torch.empty(9 * 64 * 128, device=self.dev, dtype=torch.uint8)
                _lib.check(self.lib.dfir_pack_conv3x3_bf16(wi.data_ptr(), t.data_ptr(), 64, 64, 64, 0, 1, self.st), "pack")
                prev = skip if i == 0 else out
                _lib.check(self.lib.dfir_conv3x3_c64_accumulate(chunk.data_ptr(), t.data_ptr(),
                                                                conv.bias.detach().data_ptr() if i == nc - 1 else None,
                                                                self.B, self.H, self.W, None, _ptr(prev), out.data_ptr(),
                                                                junk.data_ptr(), 0, self.st), "conv tc")
            return out
        w = torch.empty(9 * cin * cout, **self.f32)
        _lib.check(self.lib.dfir_pack_conv3x3_f32(conv.weight.detach().contiguous().data_ptr(), w.data_ptr(), cout, cin,
                                                  self.st), "pack")
        out = self.feat(cout)
        _lib.check(self.lib.dfir_conv3x3_f32(x.data_ptr(), w.data_ptr(), conv.bias.detach().data_ptr(), _ptr(skip),
                                             out.data_ptr(), self.B, self.H, self.W, cin, cout, 0, 1, 0, self.st), "conv")
        return out

    def conv_bwd(self, conv, x, dy, gw, gb, need_dx=True):
        """weight / bias gradients into gw / gb (views of the flat gradient buffer), returns dL/dx"""
        cin, cout = conv.in_channels, conv.out_channels
        if self._tc(conv):
            nc = cin // 64
            w = conv.weight.detach()
            dyb = dy.to(torch.bfloat16)
            n = self.lib.dfir_conv3x3_wgrad_scratch_bytes(self.B, self.H, self.W, 64, 64, 0)
            sc = self.scratch("wgrad_tc", n)
            junk = torch.empty(self.B, self.H, self.W, 64, device=self.dev, dtype=torch.bfloat16)
            jb = torch.empty(64, **self.f32)
            dxs = []
            for i, chunk in enumerate(self._chunks_bf16(x)):
                gwi = gw if nc == 1 else torch.empty(64, 64, 3, 3, **self.f32)
                _lib.check(self.lib.dfir_conv3x3_wgrad_c64(dyb.data_ptr(), 0, 0, 0, chunk.data_ptr(), self.B, self.H, self.W,
                                                           gwi.data_ptr(), (gb if i == 0 else jb).data_ptr(), 0, 1,
                                                           sc.data_ptr(), sc.numel(), self.st), "wgrad tc")
                if nc > 1:
                    gw[:, i * 64:(i + 1) * 64].copy_(gwi)
                if need_dx:
                    wi = w[:, i * 64:(i + 1) * 64].contiguous()
                    t = torch.empty(9 * 64 * 128, device=self.dev, dtype=torch.uint8)
                    _lib.check(self.lib.dfir_pack_conv3x3_bf16_ex(wi.data_ptr(), t.data_ptr(), 64, 64, 0, 1, 1, self.st),
                               "pack T")
                    dxi = self.feat(64)
                    _lib.check(self.lib.dfir_conv3x3_c64_dgrad(dyb.data_ptr(), 0, 0, 0, t.data_ptr(), None, None,
                                                               dxi.data_ptr(), junk.data_ptr(), self.B, self.H, self.W,
                                                               self.st), "dgrad tc")
                    dxs.append(dxi)
            if not need_dx:
                return None
            return dxs[0] if nc == 1 else torch.cat(dxs, dim=-1)
        n = self.lib.dfir_conv3x3_wgrad_scratch_bytes(self.B, self.H, self.W, cin, cout, 1)
        sc = self.scratch("wgrad", n)
        _lib.check(self.lib.dfir_conv3x3_wgrad_f32(dy.data_ptr(), x.data_ptr(), self.B, self.H, self.W, cin, cout,
                                                   gw.data_ptr(), gb.data_ptr(), sc.data_ptr(), sc.numel(), self.st),
                   "wgrad")
        if not need_dx:
            return None
        wT = torch.empty(9 * cin * cout, **self.f32)
        _lib.check(self.lib.dfir_pack_conv3x3_f32_ex(conv.weight.detach().contiguous().data_ptr(), wT.data_ptr(), cout, cin,
                                                     1, self.st), "pack T")
        dx = self.feat(cin)
        _lib.check(self.lib.dfir_conv3x3_f32(dy.data_ptr(), wT.data_ptr(), None, None, dx.data_ptr(), self.B, self.H,
                                             self.W, cout, cin, 0, 1, 0, self.st), "dgrad")
        return dx

    # ---- elementwise / reductions
    def scale_add(self, x, svec=None, add=None, alpha=1.0, out=None):
        """out = x * svec[b][c] + alpha * add"""
        out = torch.empty_like(x) if out is None else out
        Cc = x.shape[-1]
        _lib.check(self.lib.dfir_channel_scale(x.data_ptr(), _ptr(svec), _ptr(add), float(alpha), out.data_ptr(), self.B,
                                               x.numel() // (self.B * Cc), Cc, self.st), "channel_scale")
        return out

    def dot(self, a, b, out=None, total=None, accumulate=False):
        Cc = a.shape[-1]
        sc = self.scratch("dot", self.lib.dfir_channel_dot_scratch_bytes(self.B, Cc))
        _lib.check(self.lib.dfir_channel_dot(a.data_ptr(), b.data_ptr(), _ptr(out), _ptr(total), int(accumulate),
                                             sc.data_ptr(), sc.numel(), self.B, a.numel() // (self.B * Cc), Cc, self.st),
                   "channel_dot")
        return out

    # ---- non-local attention
    def nl_params(self):
        nl = self.net.non_local.non_local
        Cf = self.C
        w_tpg = torch.cat([nl.theta.weight.detach().reshape(-1, Cf), nl.phi[0].weight.detach().reshape(-1, Cf),
                           nl.g[0].weight.detach().reshape(-1, Cf)]).contiguous()
        b_tpg = torch.cat([nl.theta.bias.detach(), nl.phi[0].bias.detach(), nl.g[0].bias.detach()]).contiguous()
        return w_tpg, b_tpg, nl.W.weight.detach().reshape(Cf, -1).contiguous(), nl.W.bias.detach().contiguous()

    def nl_fwd(self, x, prm):
        out = torch.empty_like(x)
        sc = self.scratch("nl", self.lib.dfir_nonlocal_scratch_bytes(self.B, self.H, self.W))
        _lib.check(self.lib.dfir_nonlocal(x.data_ptr(), prm[0].data_ptr(), prm[1].data_ptr(), prm[2].data_ptr(),
                                          prm[3].data_ptr(), out.data_ptr(), sc.data_ptr(), self.B, self.H, self.W, self.C,
                                          self.st), "nonlocal")
        return out

    def nl_bwd(self, x, dz, prm, gprm, accumulate):
        dx = torch.empty_like(x)
        sc = self.scratch("nlb", self.lib.dfir_nonlocal_backward_scratch_bytes(self.B, self.H, self.W))
        _lib.check(self.lib.dfir_nonlocal_backward(x.data_ptr(), dz.data_ptr(), prm[0].data_ptr(), prm[1].data_ptr(),
                                                   prm[2].data_ptr(), dx.data_ptr(), gprm[0].data_ptr(), gprm[1].data_ptr(),
                                                   gprm[2].data_ptr(), gprm[3].data_ptr(), int(accumulate), sc.data_ptr(),
                                                   sc.numel(), self.B, self.H, self.W, self.C, self.st), "nonlocal bwd")
        return dx

    # ---- second-order channel attention
    def soca_fwd(self, flow, mlp, R):
        B = self.B
        cov = torch.empty(B, 64, 64, **self.f32)
        S = torch.empty(B, 64, 64, **self.f32)
        svec = torch.empty(B, 64, **self.f32)
        sc = self.scratch("cov", self.lib.dfir_covpool_scratch_bytes(B))
        _lib.check(self.lib.dfir_covpool(flow.data_ptr(), cov.data_ptr(), sc.data_ptr(), sc.numel(), B, self.H, self.W, 64, 1,
                                         self.st), "covpool")
        _lib.check(self.lib.dfir_sqrtm(cov.data_ptr(), S.data_ptr(), B, 64, 5, self.st), "sqrtm")
        _lib.check(self.lib.dfir_soca_mlp(S.data_ptr(), mlp.data_ptr(), R, svec.data_ptr(), B, self.st), "soca mlp")
        return cov, S, svec

    def soca_bwd(self, flow, cov, S, dsvec, mlp, R, gmlp):
        """returns dL/dflow through the covariance path; gmlp: flat gradient of (W1, b1, W2, b2)"""
        B = self.B
        dS = torch.empty(B, 64, 64, **self.f32)
        sc = self.scratch("socamlp", self.lib.dfir_soca_mlp_backward_scratch_bytes(B, R))
        _lib.check(self.lib.dfir_soca_mlp_backward(S.data_ptr(), dsvec.data_ptr(), mlp.data_ptr(), R, dS.data_ptr(),
                                                   gmlp.data_ptr(), sc.data_ptr(), sc.numel(), B, self.st), "soca mlp bwd")
        dcov = torch.empty(B, 64, 64, **self.f32)
        sc = self.scratch("sqrtm", self.lib.dfir_sqrtm_scratch_bytes(B, 5))
        _lib.check(self.lib.dfir_sqrtm_backward(cov.data_ptr(), dS.data_ptr(), dcov.data_ptr(), sc.data_ptr(), sc.numel(), B,
                                                64, 5, self.st), "sqrtm bwd")
        dflow = torch.empty_like(flow)
        sc = self.scratch("cov", self.lib.dfir_covpool_scratch_bytes(B))
        _lib.check(self.lib.dfir_covpool_backward(flow.data_ptr(), dcov.data_ptr(), dflow.data_ptr(), sc.data_ptr(),
                                                  sc.numel(), B, self.H, self.W, 64, 1, self.st), "covpool bwd")
        return dflow


def _grad_map(packed, which):
    return {id(p): g for p, g in zip(packed.grad_params, packed.grad_views[which])}


def _pick_flat(packed):
    """index of the flat gradient buffer that no live .grad aliases (accumulation semantics survive a missing zero_grad)"""
    first = packed.grad_params[0].grad
    return 1 if (first is not None and first.data_ptr() == packed.grad_views[0][0].data_ptr()) else 0


def _publish(packed, net, which):
    flat = packed.grad_flat[which]
    if getattr(net, "ddp_allreduce", True):
        allreduce_mean_(flat)
    acc_p, acc_g = [], []
    unused = net.unused_parameter_ids()
    for p, g in zip(packed.grad_params, packed.grad_views[which]):
        if not p.requires_grad or id(p) in unused:  # (the reference's autograd leaves .grad of unused parameters at None)
            continue
        if p.grad is None:
            p.grad = g
        else:
            acc_p.append(p.grad)
            acc_g.append(g)
    if acc_p:
        torch._foreach_add_(acc_p, acc_g)


def _flat_or_temp(views, tensors):
    """the gradient views of `tensors` as ONE flat buffer when they are adjacent in the flat gradient buffer (no padding
    between them), else a temporary plus the copy-back list"""
    vs = [views[id(t)] for t in tensors]
    off = vs[0].data_ptr()
    for v in vs:
        if v.data_ptr() != off:
            tmp = torch.empty(sum(v.numel() for v in vs), device=vs[0].device, dtype=torch.float32)
            return tmp, vs
        off += v.numel() * 4
    return vs[0], None


def _scatter(tmp, vs):
    o = 0
    for v in vs:
        v.copy_(tmp[o:o + v.numel()].view(v.shape))
        o += v.numel()


# ====================================================================================================
# Q-SAN / SAN
# ====================================================================================================
class _QsanTrain(torch.autograd.Function):
    @staticmethod
    def forward(ctx, anchor, x, attr, net, packed):
        B, _, H, W = x.shape
        with torch.cuda.device(x.device):
            packed.repack()
            op = _Ops(net, packed, B, H, W)
            gamma = float(net.gamma.detach())
            prm = op.nl_params()
            head = op.feat()
            op.stage_fwd(T_HEAD, x=x, attr=attr, feat_out=head)
            xx = op.nl_fwd(head, prm)
            residual = xx
            saved = []
            for g, grp in enumerate(net.RG):
                flow = op.feat()
                op.stage_fwd(T_GROUPS, g, g + 1, attr=attr, feat_in=xx, feat_out=flow)
                mlp = grp.soca.flat().contiguous()
                R = grp.soca.conv_du[0].out_channels
                cov, S, svec = op.soca_fwd(flow, mlp, R)
                y = op.scale_add(flow, svec=svec)
                f = op.conv_fwd(grp.conv_last, y, skip=xx)
                saved.append((xx, flow, cov, S, svec, mlp, R))
                xx = op.scale_add(f, add=residual, alpha=gamma)
            res = op.scale_add(op.nl_fwd(xx, prm), add=head, alpha=1.0)
            out = torch.empty(B, packed.out_feats, H * packed.scale, W * packed.scale, **op.f32)
            op.stage_fwd(T_TAIL, feat_in=res, out=out)
        ctx.state = (x, attr, head, residual, xx, saved, prm, gamma)
        ctx.net, ctx.packed = net, packed
        packed.ws_owner = ctx
        return out

    @staticmethod
    def backward(ctx, gout):
        net, packed = ctx.net, ctx.packed
        if packed.ws_owner is not ctx:
            raise RuntimeError("the training workspace was re-used by another forward before this backward ran")
        x, attr, head, residual, xx_last, saved, prm, gamma = ctx.state
        B, _, H, W = x.shape
        gout = gout.to(torch.float32).contiguous()
        which = _pick_flat(packed)
        _, gstruct = packed.grad_tables[which]
        gv = _grad_map(packed, which)
        with torch.cuda.device(x.device):
            packed.grad_flat[which].zero_()
            op = _Ops(net, packed, B, H, W)
            gprm = [torch.empty_like(t) for t in prm]          # (theta|phi|g) weights, biases, W weight, W bias
            g_res = op.feat()
            op.stage_bwd(gstruct, T_TAIL, gout=gout, gfeat_in=g_res)
            d_xx = op.nl_bwd(xx_last, g_res, prm, gprm, accumulate=False)
            d_residual = torch.zeros_like(d_xx)     # gradient of the share-source skip: sum over groups of gamma * d_xx
            ggamma = gv[id(net.gamma)]
            for g in reversed(range(len(net.RG))):
                grp = net.RG[g]
                xx, flow, cov, S, svec, mlp, R = saved[g]
                # xx_next = f + gamma * residual,  f = conv_last(flow * svec) + xx
                op.dot(d_xx, residual, total=ggamma, accumulate=True)   # (the flat buffer was zeroed above)
                op.scale_add(d_residual, add=d_xx, alpha=gamma, out=d_residual)
                y = op.scale_add(flow, svec=svec)
                d_y = op.conv_bwd(grp.conv_last, y, d_xx, gv[id(grp.conv_last.weight)], gv[id(grp.conv_last.bias)])
                dsvec = torch.empty(B, 64, **op.f32)
                op.dot(d_y, flow, out=dsvec)
                du = grp.soca.conv_du
                gmlp, back = _flat_or_temp(gv, [du[0].weight, du[0].bias, du[2].weight, du[2].bias])
                d_flow_cov = op.soca_bwd(flow, cov, S, dsvec, mlp, R, gmlp)
                if back is not None:
                    _scatter(gmlp, back)
                d_flow = op.scale_add(d_y, svec=svec, add=d_flow_cov, alpha=1.0)
                d_in = op.feat()
                op.stage_bwd(gstruct, T_GROUPS, g, g + 1, attr=attr, gfeat_out=d_flow, gfeat_in=d_in)
                d_xx = op.scale_add(d_in, add=d_xx, alpha=1.0)
            d_xx0 = op.scale_add(d_xx, add=d_residual, alpha=1.0)
            d_head = op.nl_bwd(head, d_xx0, prm, gprm, accumulate=True)
            d_head = op.scale_add(d_head, add=g_res, alpha=1.0)
            op.stage_bwd(gstruct, T_HEAD, x=x, gfeat_out=d_head)
            op.stage_bwd(gstruct, T_ATTN, attr=attr)
            nl = net.non_local.non_local
            for i, m in enumerate((nl.theta, nl.phi[0], nl.g[0])):
                gv[id(m.weight)].copy_(gprm[0][8 * i: 8 * i + 8].view(m.weight.shape))
                gv[id(m.bias)].copy_(gprm[1][8 * i: 8 * i + 8])
            gv[id(nl.W.weight)].copy_(gprm[2].view(nl.W.weight.shape))
            gv[id(nl.W.bias)].copy_(gprm[3])
        _publish(packed, net, which)
        ctx.state = None
        return gout.new_zeros(1), None, None, None, None


# ====================================================================================================
# Q-HAN / HAN
# ====================================================================================================
class _QhanTrain(torch.autograd.Function):
    @staticmethod
    def forward(ctx, anchor, x, attr, net, packed):
        B, _, H, W = x.shape
        ng, Cf = net.cfg["n_resgroups"], net.cfg["n_feats"]
        with torch.cuda.device(x.device):
            packed.repack()
            op = _Ops(net, packed, B, H, W)
            lib = op.lib
            head = op.feat()
            op.stage_fwd(T_HEAD, x=x, attr=attr, feat_out=head)
            stack = torch.empty(ng + 1, B, H, W, Cf, **op.f32)   # [g] = output of group g, [ng] = body.<ng> conv output
            op.stage_fwd(T_GROUPS, 0, ng, attr=attr, feat_in=head, group_out=stack)
            stack[ng] = op.conv_fwd(net.body[ng], stack[ng - 1])
            la = op.feat((ng + 1) * Cf)
            lam_sc = torch.empty(int(lib.dfir_lam_scratch_bytes(B, ng + 1)), device=op.dev, dtype=torch.uint8)
            per_map = B * H * W * Cf
            gamma_la = float(net.la.gamma.detach())
            _lib.check(lib.dfir_lam(stack[ng].data_ptr(), -per_map, gamma_la, la.data_ptr(), lam_sc.data_ptr(), ng + 1, B,
                                    H * W, Cf, op.st), "lam")
            out2 = op.conv_fwd(net.last_conv, la)
            out1 = op.feat()
            w27 = net.csa.conv.weight.detach().reshape(-1).contiguous()
            csa_b, csa_g = float(net.csa.conv.bias.detach()), float(net.csa.gamma.detach())
            _lib.check(lib.dfir_csam(stack[ng].data_ptr(), w27.data_ptr(), csa_b, csa_g, out1.data_ptr(), B, H, W, Cf,
                                     op.st), "csam")
            cat = torch.cat([out1, out2], dim=-1)
            res = op.conv_fwd(net.last, cat, skip=head)
            out = torch.empty(B, packed.out_feats, H * packed.scale, W * packed.scale, **op.f32)
            op.stage_fwd(T_TAIL, feat_in=res, out=out)
        ctx.state = (x, attr, stack, la, lam_sc, cat, w27, csa_b, csa_g, gamma_la)
        ctx.net, ctx.packed = net, packed
        packed.ws_owner = ctx
        return out

    @staticmethod
    def backward(ctx, gout):
        net, packed = ctx.net, ctx.packed
        if packed.ws_owner is not ctx:
            raise RuntimeError("the training workspace was re-used by another forward before this backward ran")
        x, attr, stack, la, lam_sc, cat, w27, csa_b, csa_g, gamma_la = ctx.state
        B, _, H, W = x.shape
        ng, Cf = net.cfg["n_resgroups"], net.cfg["n_feats"]
        gout = gout.to(torch.float32).contiguous()
        which = _pick_flat(packed)
        _, gstruct = packed.grad_tables[which]
        gv = _grad_map(packed, which)
        with torch.cuda.device(x.device):
            packed.grad_flat[which].zero_()
            op = _Ops(net, packed, B, H, W)
            lib = op.lib
            g_res = op.feat()
            op.stage_bwd(gstruct, T_TAIL, gout=gout, gfeat_in=g_res)
            d_cat = op.conv_bwd(net.last, cat, g_res, gv[id(net.last.weight)], gv[id(net.last.bias)])
            d_out1 = d_cat[..., :Cf].contiguous()
            d_out2 = d_cat[..., Cf:].contiguous()
            # channel-spatial attention on the body output
            d_top = op.feat()
            sc = op.scratch("csam", lib.dfir_csam_backward_scratch_bytes(B, H, W, Cf))
            _lib.check(lib.dfir_csam_backward(stack[ng].data_ptr(), d_out1.data_ptr(), w27.data_ptr(), csa_b, csa_g,
                                              d_top.data_ptr(), gv[id(net.csa.conv.weight)].data_ptr(),
                                              gv[id(net.csa.conv.bias)].data_ptr(), gv[id(net.csa.gamma)].data_ptr(),
                                              sc.data_ptr(), sc.numel(), B, H, W, Cf, op.st), "csam bwd")
            # fusion conv over the layer-attention maps, then LAM
            d_la = op.conv_bwd(net.last_conv, la, d_out2, gv[id(net.last_conv.weight)], gv[id(net.last_conv.bias)])
            dstack = torch.empty_like(stack)
            per_map = B * H * W * Cf
            sc = op.scratch("lam", lib.dfir_lam_backward_scratch_bytes(B, ng + 1))
            _lib.check(lib.dfir_lam_backward(stack[ng].data_ptr(), -per_map, lam_sc.data_ptr(), gamma_la, d_la.data_ptr(),
                                             dstack[ng].data_ptr(), -per_map, gv[id(net.la.gamma)].data_ptr(),
                                             sc.data_ptr(), sc.numel(), ng + 1, B, H * W, Cf, op.st), "lam bwd")
            d_top = op.scale_add(d_top, add=dstack[ng], alpha=1.0)
            tail_conv = net.body[ng]
            d_prev = op.conv_bwd(tail_conv, stack[ng - 1], d_top, gv[id(tail_conv.weight)], gv[id(tail_conv.bias)])
            d_cur = op.scale_add(d_prev, add=dstack[ng - 1], alpha=1.0)
            for g in reversed(range(ng)):
                d_in = op.feat()
                op.stage_bwd(gstruct, T_GROUPS, g, g + 1, attr=attr, gfeat_out=d_cur, gfeat_in=d_in)
                d_cur = op.scale_add(d_in, add=dstack[g - 1], alpha=1.0) if g > 0 else d_in
            d_head = op.scale_add(d_cur, add=g_res, alpha=1.0)
            op.stage_bwd(gstruct, T_HEAD, x=x, gfeat_out=d_head)
            op.stage_bwd(gstruct, T_ATTN, attr=attr)
        _publish(packed, net, which)
        ctx.state = None
        return gout.new_zeros(1), None, None, None, None


def staged_train_apply(net, packed, x, attr, kind):
    if x.requires_grad or attr.requires_grad:
        raise RuntimeError("the B200 training path does not compute gradients with respect to the input image or the "
                           "metadata: detach them (x.requires_grad=%s, metadata.requires_grad=%s)"
                           % (x.requires_grad, attr.requires_grad))
    anchor = getattr(net, "_train_anchor", None)
    if anchor is None or anchor.device != x.device:
        anchor = torch.zeros(1, device=x.device, requires_grad=True)
        net._train_anchor = anchor
    fn = _QsanTrain if kind == "san" else _QhanTrain
    return fn.apply(anchor, x, attr, net, packed)
