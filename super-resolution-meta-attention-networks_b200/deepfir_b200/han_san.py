"""Q-HAN and Q-SAN on the B200 path.

Both reuse the Q-RCAN machinery for their 64-channel conv trunks (`dfir_qrcan_stages`: head conv, residual
groups on the tensor cores, upsampler + tail) and add their own layers through dedicated kernels
(csrc/san_han.cu): layer attention (LAM), channel-spatial attention (CSAM), second-order channel attention
(covariance pooling + Newton-Schulz square root) and region non-local attention.  Parameter names, shapes and
registration order mirror the reference (`attention_manipulators/architectures.py:402-540`,
`attention_manipulators/qsan_blocks.py`, `advanced/SAN_blocks.py`, `advanced/HAN_blocks.py`).
"""
import ctypes as C

import torch
from torch import nn

from . import _lib
from .qrcan import (PRECISIONS, SCHEDULES, MetaAttentionParams, PackedQrcan, QResidualGroupParams, UpsamplerParams,
                    _conv3, _fc)


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


class _StagedNet(nn.Module):
    """shared plumbing: packed trunk, workspace, single-op wrappers (all NHWC fp32 device tensors)."""

    precision = "bf16"
    schedule = "linear"
    chunk_images = 1 << 20  # staged execution keeps its state in the workspace: always one pass
    _packed = None

    def packed(self, training=False):
        """Kernel-format parameters of the conv trunk (see QRCAN.packed): rebuilt when a parameter's storage moved,
        refreshed in place by `dfir_qrcan_repack` when only values changed (optimizer.step()); the cached kernel-format
        copies of the layers outside the trunk (`_side`) are dropped with either."""
        params = self.__dict__.get("_plist")
        if params is None:
            params = self.__dict__["_plist"] = list(self.parameters())
        skey = (tuple(p.data_ptr() for p in params), self.precision, self.schedule)
        vers = tuple(p._version for p in params)
        pk = self._packed
        if pk is None or pk.key != skey or (pk.versions != vers and not pk.can_repack):
            if pk is not None:
                pk.close()
            pk = self._packed = PackedQrcan(self, skey)
            pk.versions = vers
            self._side = {}
        if training and not pk.train_ready:
            pk.enable_training(self)
            pk.versions = None
        if training:
            pk.versions = None          # the training forward re-packs the parameters itself
        elif pk.versions != vers:
            pk.repack()
            pk.versions = vers
            self._side = {}
        return pk

    def _apply(self, fn, *a, **k):
        self._packed = None
        self.__dict__.pop("_plist", None)
        return super()._apply(fn, *a, **k)

    def unused_parameter_ids(self):
        """parameters the reference constructs and serialises but never calls in forward: their .grad stays None"""
        return set()

    def _training_step(self, x):
        return torch.is_grad_enabled() and self.head[0].weight.requires_grad

    def _train_forward(self, x, metadata, kind):
        from .train_staged import staged_train_apply
        pk = self.packed(training=True)
        B = x.shape[0]
        attr = metadata.reshape(B, -1).to(device=x.device, dtype=torch.float32).contiguous()
        return staged_train_apply(self, pk, x.to(torch.float32).contiguous(), attr, kind)

    def invalidate_packed(self):
        """rebuild the kernel-format parameters on the next forward (see QRCAN.invalidate_packed: needed after in-place
        writes through `.data`, which `Parameter._version` does not see)"""
        if self._packed is not None:
            self._packed.close()
        self._packed = None

    def train(self, mode=True):
        if self.training and not mode:
            self.invalidate_packed()
        return super().train(mode)

    def load_state_dict(self, *a, **k):
        out = super().load_state_dict(*a, **k)
        self.invalidate_packed()
        self.__dict__.pop("_plist", None)
        return out

    # -- staged trunk -------------------------------------------------------------------------------
    def _stages(self, pk, stages, g0, g1, B, H, W, x=None, attr=None, feat_in=None, group_out=None, feat_out=None,
                out=None):
        lib = _lib.load_library()
        prec = PRECISIONS[self.precision]
        ws = pk.workspace(B, H, W, prec)
        ptr = lambda t: (t.data_ptr() if t is not None else None)
        rc = lib.dfir_qrcan_stages(C.byref(pk.desc), stages, g0, g1, ptr(x), ptr(attr), ptr(feat_in), ptr(group_out),
                                   ptr(feat_out), ptr(out), B, H, W, prec, ws.data_ptr(), ws.numel(),
                                   _stream(ws.device))
        _lib.check(rc, "qrcan_stages(%d)" % stages)

    # -- single ops -----------------------------------------------------------------------------------
    def _packed_f32(self, name, conv):
        w = self._side.get(name)
        if w is None:
            lib = _lib.load_library()
            wt = conv.weight.detach().contiguous()
            w = torch.empty(9 * wt.shape[1] * wt.shape[0], device=wt.device, dtype=torch.float32)
            _lib.check(lib.dfir_pack_conv3x3_f32(wt.data_ptr(), w.data_ptr(), wt.shape[0], wt.shape[1],
                                                 _stream(wt.device)), "pack %s" % (name,))
            self._side[name] = w
        return w

    def _conv(self, name, conv, x, skip=None):
        """3x3 conv of an fp32 NHWC tensor outside the staged trunk (group / fusion convs of Q-SAN and Q-HAN).  bf16 mode
        with 64 output channels and Cin a multiple of 64: the tensor-core kernel, one launch per 64-channel input chunk
        with the fp32 running sum chained through its skip input (`dfir_conv3x3_c64_accumulate`); else CUDA cores."""
        Cin = x.shape[-1]
        if self.precision != "bf16" or conv.out_channels != 64 or Cin % 64 != 0:
            return self._conv_f32(name, conv, x, skip=skip)
        lib = _lib.load_library()
        B, H, W, _ = x.shape
        nc = Cin // 64
        tiles = self._side.get(("tc", name))
        if tiles is None:
            tiles = []
            wt = conv.weight.detach().float()
            for i in range(nc):
                wi = wt[:, i * 64:(i + 1) * 64].contiguous()
                t = torch.empty(9 * 64 * 128, device=wt.device, dtype=torch.uint8)
                _lib.check(lib.dfir_pack_conv3x3_bf16(wi.data_ptr(), t.data_ptr(), 64, 64, 64, 0, 1, _stream(wt.device)),
                           "pack %s" % (name,))
                tiles.append(t)
            tiles.append(conv.bias.detach().float().contiguous())
            self._side[("tc", name)] = tiles
        out = torch.empty(B, H, W, 64, device=x.device, dtype=torch.float32)
        junk = torch.empty(B, H, W, 64, device=x.device, dtype=torch.bfloat16)
        xb = x.to(torch.bfloat16)
        for i in range(nc):
            chunk = xb if nc == 1 else xb[..., i * 64:(i + 1) * 64].contiguous()
            prev = skip if i == 0 else out
            _lib.check(lib.dfir_conv3x3_c64_accumulate(chunk.data_ptr(), tiles[i].data_ptr(),
                                                       tiles[nc].data_ptr() if i == nc - 1 else None, B, H, W, None,
                                                       prev.data_ptr() if prev is not None else None, out.data_ptr(),
                                                       junk.data_ptr(), 0, _stream(x.device)), "conv %s" % (name,))
        return out

    def _conv_f32(self, name, conv, x, skip=None):
        lib = _lib.load_library()
        B, H, W, Cin = x.shape
        out = torch.empty(B, H, W, conv.out_channels, device=x.device, dtype=torch.float32)
        w = self._packed_f32(name, conv)
        _lib.check(lib.dfir_conv3x3_f32(x.data_ptr(), w.data_ptr(), conv.bias.detach().data_ptr(),
                                        skip.data_ptr() if skip is not None else None, out.data_ptr(), B, H, W, Cin,
                                        conv.out_channels, 0, 1, 0, _stream(x.device)), "conv %s" % (name,))
        return out

    def _scale_add(self, x, svec=None, add=None, alpha=1.0):
        lib = _lib.load_library()
        B, H, W, Cc = x.shape
        out = torch.empty_like(x)
        _lib.check(lib.dfir_channel_scale(x.data_ptr(), svec.data_ptr() if svec is not None else None,
                                          add.data_ptr() if add is not None else None, float(alpha), out.data_ptr(), B,
                                          H * W, Cc, _stream(x.device)), "channel_scale")
        return out

    def _check_input(self, x):
        if not x.is_cuda:
            raise RuntimeError("deepfir_b200.%s runs on a CUDA (sm_100a) device only: there is no CPU path"
                               % type(self).__name__)

    def forensic(self, *args, **kwargs):
        raise NotImplementedError("forensic analysis is outside the B200 hot path")


# ====================================================================================================
# Q-HAN
# ====================================================================================================
class LAMParams(nn.Module):
    def __init__(self):
        super().__init__()
        self.gamma = nn.Parameter(torch.zeros(1))


class CSAMParams(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv = nn.Conv3d(1, 1, 3, 1, 1)
        self.gamma = nn.Parameter(torch.zeros(1))


class QHAN(_StagedNet):
    """reference attention_manipulators/architectures.py:470-540 (n_resgroups must be 10: `last_conv` is built
    for 11 stacked feature maps)."""

    def __init__(self, n_resgroups=10, n_resblocks=20, n_feats=64, reduction=16, num_metadata=0, scale=4, n_colors=3,
                 res_scale=1.0, num_q_layers_inner_residual=None, precision='bf16', schedule='linear', **kwargs):
        super().__init__()
        if precision not in PRECISIONS or schedule not in SCHEDULES:
            raise RuntimeError("unknown precision / schedule")
        self.precision, self.schedule, self.scale, self.style = precision, schedule, scale, "standard"
        M = num_metadata
        self.cfg = dict(n_resblocks=n_resblocks, n_resgroups=n_resgroups, n_feats=n_feats, in_feats=n_colors,
                        out_feats=n_colors, scale=scale, reduction=reduction, num_metadata=M, style="standard",
                        no_group_conv=0, meta_relu=1, res_scale=1.0,
                        meta_hidden=(n_feats // 2 if M <= 15 else (n_feats - M) // 2 + M))
        head = [_conv3(n_colors, n_feats)]
        body = [QResidualGroupParams(n_feats, reduction, n_resblocks, "standard", M, False, True,
                                     num_q_layers_inner_residual) for _ in range(n_resgroups)]
        body.append(_conv3(n_feats, n_feats))
        tail = [UpsamplerParams(scale, n_feats), _conv3(n_feats, n_colors)]
        self.head = nn.Sequential(*head)
        self.body = nn.Sequential(*body)
        self.csa = CSAMParams()
        self.la = LAMParams()
        self.last_conv = nn.Conv2d(n_feats * 11, n_feats, 3, 1, 1)
        self.last = nn.Conv2d(n_feats * 2, n_feats, 3, 1, 1)
        self.tail = nn.Sequential(*tail)

    def _pack_spec(self):
        cfg = self.cfg
        ng, nb = cfg["n_resgroups"], cfg["n_resblocks"]
        trunk = []
        for g in range(ng):
            grp = self.body[g]
            for b in range(nb):
                trunk += [grp.body[b].body[0], grp.body[b].body[2]]
            trunk.append(grp.final_body)
        trunk.append(self.body[ng])  # slot of the trunk tail conv (run as a separate op: HAN adds no head skip here)
        blocks = [self.body[g].body[b] for g in range(ng) for b in range(nb)]
        return dict(cfg=cfg, head=self.head[0], trunk=trunk,
                    ups=[m for m in self.tail[0] if isinstance(m, nn.Conv2d)], tail=self.tail[1],
                    ca=[blk.final_body.flat_params() for blk in blocks],
                    ca_params=[blk.final_body.param_list() for blk in blocks],
                    meta=[tuple(blk.q_node.fcs()) if blk.q_layer else None for blk in blocks])

    def forward(self, x, metadata):
        self._check_input(x)
        if self._training_step(x):  # forward with saved activations; backward fills every .grad (train_staged.py)
            return self._train_forward(x, metadata, "han")
        lib = _lib.load_library()
        pk = self.packed()
        ng, Cf = self.cfg["n_resgroups"], self.cfg["n_feats"]
        B, _, H, W = x.shape
        dev = x.device
        x = x.to(torch.float32).contiguous()
        attr = metadata.reshape(B, -1).to(device=dev, dtype=torch.float32).contiguous()
        f32 = dict(device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            head = torch.empty(B, H, W, Cf, **f32)
            self._stages(pk, 1, 0, 0, B, H, W, x=x, feat_out=head)
            stack = torch.empty(ng + 1, B, H, W, Cf, **f32)  # [g] = output of group g, [ng] = body.<ng> conv output
            self._stages(pk, 2, 0, ng, B, H, W, attr=attr, group_out=stack)
            stack[ng] = self._conv("body_tail", self.body[ng], stack[ng - 1])
            # LAM over the maps in the reference's order (newest first): map n = stack[ng - n]
            la = torch.empty(B, H, W, (ng + 1) * Cf, **f32)
            scratch = torch.empty(int(lib.dfir_lam_scratch_bytes(B, ng + 1)), device=dev, dtype=torch.uint8)
            per_map = B * H * W * Cf
            _lib.check(lib.dfir_lam(stack[ng].data_ptr(), -per_map, float(self.la.gamma), la.data_ptr(),
                                    scratch.data_ptr(), ng + 1, B, H * W, Cf, _stream(dev)), "lam")
            out2 = self._conv("last_conv", self.last_conv, la)
            out1 = torch.empty(B, H, W, Cf, **f32)
            _lib.check(lib.dfir_csam(stack[ng].data_ptr(), self.csa.conv.weight.detach().reshape(-1).data_ptr(),
                                     float(self.csa.conv.bias), float(self.csa.gamma), out1.data_ptr(), B, H, W, Cf,
                                     _stream(dev)), "csam")
            res = self._conv("last", self.last, torch.cat([out1, out2], dim=-1), skip=head)
            out = torch.empty(B, self.cfg["out_feats"], H * self.scale, W * self.scale, **f32)
            self._stages(pk, 8, 0, 0, B, H, W, feat_in=res, out=out)
        return out


# ====================================================================================================
# Q-SAN
# ====================================================================================================
class QRBParams(nn.Module):
    def __init__(self, n_feat, num_metadata):
        super().__init__()
        self.conv_first = nn.Sequential(_conv3(n_feat, n_feat), nn.ReLU(inplace=True), _conv3(n_feat, n_feat))
        self.q_layer = MetaAttentionParams(n_feat, num_metadata, nonlinearity=True, num_layers=2)


class SOCAParams(nn.Module):
    def __init__(self, channel, reduction=8):
        super().__init__()
        self.conv_du = nn.Sequential(_fc(channel, channel // reduction), nn.ReLU(inplace=True),
                                     _fc(channel // reduction, channel), nn.Sigmoid())

    def flat(self):
        return torch.cat([t.detach().reshape(-1) for t in (self.conv_du[0].weight, self.conv_du[0].bias,
                                                           self.conv_du[2].weight, self.conv_du[2].bias)])


class QLSRAGParams(nn.Module):
    def __init__(self, n_feat, reduction, n_resblocks, num_metadata):
        super().__init__()
        self.rcab = nn.ModuleList([QRBParams(n_feat, num_metadata) for _ in range(n_resblocks)])
        self.soca = SOCAParams(n_feat, reduction=reduction)
        self.conv_last = _conv3(n_feat, n_feat)
        self.gamma = nn.Parameter(torch.zeros(1))  # serialised but unused by the reference's forward


class NonLocalBlockParams(nn.Module):
    """_NonLocalBlockND (2-D, embedded gaussian, no BN): g and phi are wrapped in Sequential(conv, MaxPool2d(2))
    because `sub_sample` is shadowed by nn.Upsample in the reference (SAN_blocks.py:36-40, 88-93)."""

    def __init__(self, in_channels, inter_channels):
        super().__init__()
        g = _fc(in_channels, inter_channels)
        self.W = _fc(inter_channels, in_channels)
        nn.init.constant_(self.W.weight, 0)
        nn.init.constant_(self.W.bias, 0)
        self.theta = _fc(in_channels, inter_channels)
        phi = _fc(in_channels, inter_channels)
        self.g = nn.Sequential(g, nn.MaxPool2d(kernel_size=2))
        self.phi = nn.Sequential(phi, nn.MaxPool2d(kernel_size=2))
        for k in ("g", "W", "theta", "phi"):  # registration order of the reference
            self._modules[k] = self._modules.pop(k)


class NonlocalCAParams(nn.Module):
    def __init__(self, in_feat=64, inter_feat=8, reduction=8):
        super().__init__()
        self.soca = SOCAParams(in_feat, reduction=reduction)  # constructed, serialised, never called
        self.non_local = NonLocalBlockParams(in_feat, inter_feat)


class QSAN(_StagedNet):
    """reference attention_manipulators/architectures.py:402-467."""

    def __init__(self, n_resgroups=20, n_resblocks=10, n_feats=64, reduction=16, scale=4, rgb_range=255, n_colors=3,
                 res_scale=1, input_para=1, precision='bf16', schedule='linear', **kwargs):
        super().__init__()
        if precision not in PRECISIONS or schedule not in SCHEDULES:
            raise RuntimeError("unknown precision / schedule")
        self.precision, self.schedule, self.scale, self.style = precision, schedule, scale, "none"
        M = input_para
        self.cfg = dict(n_resblocks=n_resblocks, n_resgroups=n_resgroups, n_feats=n_feats, in_feats=n_colors,
                        out_feats=n_colors, scale=scale, reduction=reduction, num_metadata=M, style="none",
                        no_group_conv=1, meta_relu=1, res_scale=1.0,
                        meta_hidden=(n_feats // 2 if M <= 15 else (n_feats - M) // 2 + M))
        head = [_conv3(n_colors, n_feats)]
        self.gamma = nn.Parameter(torch.zeros(1))
        self.RG = nn.ModuleList([QLSRAGParams(n_feats, reduction, n_resblocks, M) for _ in range(n_resgroups)])
        self.conv_last = _conv3(n_feats, n_feats)  # serialised but unused by the reference's forward
        tail = [UpsamplerParams(scale, n_feats), _conv3(n_feats, n_colors)]
        self.non_local = NonlocalCAParams(in_feat=n_feats, inter_feat=n_feats // 8, reduction=8)
        self.head = nn.Sequential(*head)
        self.tail = nn.Sequential(*tail)

    def _pack_spec(self):
        trunk = []
        for grp in self.RG:
            for blk in grp.rcab:
                trunk += [blk.conv_first[0], blk.conv_first[2]]
        trunk.append(self.conv_last)  # slot of the trunk tail conv (never run: the groups are driven one by one)
        blocks = [blk for grp in self.RG for blk in grp.rcab]
        return dict(cfg=self.cfg, head=self.head[0], trunk=trunk,
                    ups=[m for m in self.tail[0] if isinstance(m, nn.Conv2d)], tail=self.tail[1],
                    ca=[None for _ in blocks], ca_params=[None for _ in blocks],
                    meta=[tuple(blk.q_layer.fcs()) if hasattr(blk, "q_layer") else None for blk in blocks])

    def unused_parameter_ids(self):
        mods = [self.conv_last, self.non_local.soca]
        ids = {id(p) for m in mods for p in m.parameters()}
        return ids | {id(grp.gamma) for grp in self.RG}

    def _nonlocal(self, x):
        lib = _lib.load_library()
        B, H, W, Cf = x.shape
        nl = self.non_local.non_local
        side = self._side.get("nl")
        if side is None:
            w_tpg = torch.cat([nl.theta.weight.detach().reshape(-1, Cf), nl.phi[0].weight.detach().reshape(-1, Cf),
                               nl.g[0].weight.detach().reshape(-1, Cf)]).contiguous()
            b_tpg = torch.cat([nl.theta.bias.detach(), nl.phi[0].bias.detach(), nl.g[0].bias.detach()]).contiguous()
            side = (w_tpg, b_tpg, nl.W.weight.detach().reshape(Cf, -1).contiguous(), nl.W.bias.detach().contiguous())
            self._side["nl"] = side
        out = torch.empty_like(x)
        scratch = torch.empty(int(lib.dfir_nonlocal_scratch_bytes(B, H, W)), device=x.device, dtype=torch.uint8)
        _lib.check(lib.dfir_nonlocal(x.data_ptr(), side[0].data_ptr(), side[1].data_ptr(), side[2].data_ptr(),
                                     side[3].data_ptr(), out.data_ptr(), scratch.data_ptr(), B, H, W, Cf,
                                     _stream(x.device)), "nonlocal")
        return out

    def forward(self, x, metadata):
        self._check_input(x)
        if self._training_step(x):  # forward with saved activations; backward fills every .grad (train_staged.py)
            return self._train_forward(x, metadata, "san")
        lib = _lib.load_library()
        pk = self.packed()
        Cf = self.cfg["n_feats"]
        B, _, H, W = x.shape
        dev = x.device
        x = x.to(torch.float32).contiguous()
        attr = metadata.reshape(B, -1).to(device=dev, dtype=torch.float32).contiguous()
        f32 = dict(device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            head = torch.empty(B, H, W, Cf, **f32)
            self._stages(pk, 1, 0, 0, B, H, W, x=x, feat_out=head)
            xx = self._nonlocal(head)
            residual = xx
            gamma = float(self.gamma)
            soca_scratch = torch.empty(int(lib.dfir_soca_scratch_bytes(B)), device=dev, dtype=torch.uint8)
            svec = torch.empty(B, Cf, **f32)
            flow = torch.empty(B, H, W, Cf, **f32)
            for g, grp in enumerate(self.RG):
                # 10 x QRB: conv-ReLU-conv, meta-attention scale, + x  (tensor-core trunk, style NONE)
                self._stages(pk, 2, g, g + 1, B, H, W, attr=attr, feat_in=xx, feat_out=flow)
                mlp = self._side.get(("soca", g))
                if mlp is None:
                    mlp = grp.soca.flat().contiguous()
                    self._side[("soca", g)] = mlp
                _lib.check(lib.dfir_soca(flow.data_ptr(), mlp.data_ptr(), grp.soca.conv_du[0].out_channels,
                                         svec.data_ptr(), soca_scratch.data_ptr(), B, H, W, Cf, _stream(dev)), "soca")
                y = self._scale_add(flow, svec=svec)
                f = self._conv(("conv_last", g), grp.conv_last, y, skip=xx)       # + group input
                xx = self._scale_add(f, add=residual, alpha=gamma)                  # + gamma * share-source skip
            res = self._scale_add(self._nonlocal(xx), add=head, alpha=1.0)
            out = torch.empty(B, self.cfg["out_feats"], H * self.scale, W * self.scale, **f32)
            self._stages(pk, 8, 0, 0, B, H, W, feat_in=res, out=out)
        return out
