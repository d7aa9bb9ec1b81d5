"""Q-RCAN on the B200 path.

``QRCAN`` keeps the reference's constructor signature, parameter names, shapes (OIHW fp32) and parameter
registration order (``/root/reference/Code/SISR/models/attention_manipulators/architectures.py:246-316``)
so checkpoints, optimizers and ``print_parameters`` are interchangeable — but its modules are only
parameter containers.  ``forward`` hands the whole network to ``libdfir_b200.so`` (one C call that
enqueues every kernel on the current CUDA stream).  The kernel-format weights (bf16 swizzled tiles,
fp32 tap-major copies, attention blobs) are a derived cache keyed on the parameters' version counters;
they are rebuilt after ``optimizer.step()`` / ``load_state_dict`` and never serialised.
"""
import ctypes as C
import math

import torch
from torch import nn

from . import _lib
from ._lib import QrcanNet

STYLES = {"none": 0, "standard": 1, "modulate": 2, "max_concat": 3, "softmax": 4, "mini_concat": 5, "extended_attention": 6}
PRECISIONS = {"bf16": 0, "fp32": 1}
# block-chain schedules of the bf16 path (csrc/api.cu): pool-by-linearity, fused-in, streamer
SCHEDULES = {"linear": 0, "fused": 1, "streamer": 2, "linear3": 3}


def _conv3(cin, cout):
    return nn.Conv2d(cin, cout, 3, padding=1, bias=True)


def _fc(cin, cout):
    return nn.Conv2d(cin, cout, 1, padding=0, bias=True)


class MetaAttentionParams(nn.Module):
    """Parameter container of the meta-attention layer (q_layer.py:4-43): FC stack on the metadata."""

    def __init__(self, channels, num_metadata, nonlinearity=False, num_layers=2):
        super().__init__()
        widths = [num_metadata]
        stack = []
        for left in range(num_layers, 0, -1):
            nxt = (channels - num_metadata) // left + num_metadata if num_metadata > 15 else channels // left
            stack.append(_fc(widths[-1], nxt))
            widths.append(nxt)
            if nonlinearity and left != 1:
                stack.append(nn.ReLU(inplace=True))
        stack.append(nn.Sigmoid())
        self.attribute_integrator = nn.Sequential(*stack)
        self.widths = widths
        self.nonlinearity = nonlinearity

    def fcs(self):
        return [m for m in self.attribute_integrator if isinstance(m, nn.Conv2d)]


class ChannelAttentionParams(nn.Module):
    """Parameter container of QCALayer (architectures.py:34-103), all six styles."""

    def __init__(self, channel, style, reduction=16, num_metadata=1):
        super().__init__()
        if reduction < 16:
            raise RuntimeError('Using an extreme channel attention reduction value')
        if style not in STYLES:
            raise NotImplementedError
        self.style = style
        red = channel // reduction
        cin = channel if style in ("modulate", "mini_concat", "standard") else channel + num_metadata
        if style in ("modulate", "max_concat", "softmax", "standard"):
            self.conv_du = nn.Sequential(_fc(cin, red), nn.ReLU(inplace=True), _fc(red, channel), nn.Sigmoid())
        elif style == "mini_concat":
            self.pre_concat = _fc(cin, red)
            self.conv_du = nn.Sequential(nn.ReLU(inplace=True), _fc(red + num_metadata, channel), nn.Sigmoid())
        else:  # extended_attention
            plan = [(cin, channel // 2), (channel // 2 + num_metadata, channel // 4),
                    (channel // 4 + num_metadata, red)]
            self.feature_convs = nn.ModuleList(
                [nn.Sequential(_fc(i, o), nn.ReLU(inplace=True)) for i, o in plan])
            self.final_conv = nn.Sequential(_fc(red, channel), nn.Sigmoid())

    def param_list(self):
        """the nn.Parameters in the order the kernels expect (csrc/attn.cuh)"""
        if self.style in ("modulate", "max_concat", "softmax", "standard"):
            mods = [self.conv_du[0], self.conv_du[2]]
        elif self.style == "mini_concat":
            mods = [self.pre_concat, self.conv_du[1]]
        else:
            mods = [s[0] for s in self.feature_convs] + [self.final_conv[0]]
        out = []
        for m in mods:
            out += [m.weight, m.bias]
        return out

    def flat_params(self):
        """fp32 arrays in the order the kernels expect (csrc/simt.cu attn_vector)."""
        return [p.reshape(-1) for p in self.param_list()]


class PixelAttentionParams(nn.Module):
    def __init__(self, channel):
        super().__init__()
        self.pa = nn.Sequential(_fc(channel, channel // 8), nn.ReLU(inplace=True), _fc(channel // 8, 1), nn.Sigmoid())


class QRCABParams(nn.Module):
    """QRCAB (architectures.py:145-180).  Attribute assignment order = reference registration order."""

    def __init__(self, n_feat, reduction, style, pa, q_layer, num_metadata):
        super().__init__()
        convs = [_conv3(n_feat, n_feat), nn.ReLU(True), _conv3(n_feat, n_feat)]
        self.final_body = ChannelAttentionParams(n_feat, style, reduction, num_metadata)
        self.pa = pa
        self.q_layer = q_layer
        if pa:
            self.pa_node = PixelAttentionParams(n_feat)
        if q_layer:
            self.q_node = MetaAttentionParams(n_feat, num_metadata, nonlinearity=True)
        self.body = nn.Sequential(*convs)


class QResidualGroupParams(nn.Module):
    def __init__(self, n_feat, reduction, n_resblocks, style, num_metadata, pa, q_layer, num_q_layers):
        super().__init__()
        blocks = [QRCABParams(n_feat, reduction, style, pa,
                              q_layer if (num_q_layers is None or i < num_q_layers) else False, num_metadata)
                  for i in range(n_resblocks)]
        self.final_body = _conv3(n_feat, n_feat)
        self.body = nn.Sequential(*blocks)


class UpsamplerParams(nn.Sequential):
    """Upsampler (advanced/common.py:20-45): conv C->r^2 C + PixelShuffle(r), repeated."""

    def __init__(self, scale, n_feat):
        mods = []
        if scale & (scale - 1) == 0:
            for _ in range(int(math.log(scale, 2))):
                mods += [_conv3(n_feat, 4 * n_feat), nn.PixelShuffle(2)]
        elif scale == 3:
            mods += [_conv3(n_feat, 9 * n_feat), nn.PixelShuffle(3)]
        else:
            raise NotImplementedError
        super().__init__(*mods)


class QRCAN(nn.Module):
    def __init__(self, n_resblocks=20, n_resgroups=10, n_feats=64, in_feats=3, out_feats=3, scale=4, reduction=16,
                 res_scale=1.0, style='modulate', num_metadata=1, include_pixel_attention=False,
                 selective_meta_blocks=None, num_q_layers_inner_residual=None, include_q_layer=False,
                 precision='bf16', chunk_images=0, schedule='linear', **kwargs):
        super().__init__()
        if precision not in PRECISIONS:
            raise RuntimeError("precision must be 'bf16' or 'fp32'")
        self.style = style
        self.scale = scale
        self.precision = precision
        self.chunk_images = chunk_images
        if schedule not in SCHEDULES:
            raise RuntimeError("schedule must be one of %s" % sorted(SCHEDULES))
        self.schedule = schedule
        self.cfg = dict(n_resblocks=n_resblocks, n_resgroups=n_resgroups, n_feats=n_feats, in_feats=in_feats,
                        out_feats=out_feats, scale=scale, reduction=reduction, num_metadata=num_metadata,
                        include_pixel_attention=include_pixel_attention)
        head = [_conv3(in_feats, n_feats)]
        groups = []
        for g in range(n_resgroups):
            q = include_q_layer if (selective_meta_blocks is None or selective_meta_blocks[g]) else False
            groups.append(QResidualGroupParams(n_feats, reduction, n_resblocks, style, num_metadata,
                                               include_pixel_attention, q, num_q_layers_inner_residual))
        self.final_body = _conv3(n_feats, n_feats)
        tail = [UpsamplerParams(scale, n_feats), _conv3(n_feats, out_feats)]
        self.head = nn.Sequential(*head)
        self.body = nn.Sequential(*groups)
        self.tail = nn.Sequential(*tail)
        self._packed = None

    # ------------------------------------------------------------------ forward
    def forward(self, x, metadata):
        if not x.is_cuda:
            raise RuntimeError("deepfir_b200.%s runs on a CUDA (sm_100a) device only: there is no CPU path"
                               % type(self).__name__)
        if x.shape[0] == 0 or x.shape[2] == 0 or x.shape[3] == 0:  # empty batch / image: nothing to launch
            return x.new_zeros(x.shape[0], self.cfg["out_feats"], x.shape[2] * self.scale, x.shape[3] * self.scale,
                               dtype=torch.float32)
        from . import ops  # registers torch.ops.dfir.*
        training = torch.is_grad_enabled() and self.head_weight().requires_grad
        packed = self.packed(training=training)
        b = x.shape[0]
        attr = metadata.reshape(b, -1).to(device=x.device, dtype=torch.float32).contiguous()
        if attr.shape[1] != packed.attr_size:
            raise RuntimeError("metadata has %d entries per image, network expects %d" % (attr.shape[1], packed.attr_size))
        if training:  # forward with saved activations; backward fills every parameter's .grad (deepfir_b200/train.py)
            from .train import qrcan_train_apply
            return qrcan_train_apply(self, packed, x.to(torch.float32).contiguous(), attr)
        return torch.ops.dfir.qrcan_forward(x.to(torch.float32).contiguous(), attr, packed.handle,
                                            PRECISIONS[self.precision])

    def head_weight(self):
        head = self.head
        return (head[0] if isinstance(head, nn.Sequential) else head).weight

    def forensic(self, *args, **kwargs):
        raise NotImplementedError("forensic analysis is outside the B200 hot path")

    # ------------------------------------------------------------------ packing
    def _pack_spec(self):
        cfg = self.cfg
        ng, nb, C_, M = cfg["n_resgroups"], cfg["n_resblocks"], cfg["n_feats"], cfg["num_metadata"]
        trunk = []
        for g in range(ng):
            grp = self.body[g]
            for b in range(nb):
                trunk += [grp.body[b].body[0], grp.body[b].body[2]]
            trunk.append(grp.final_body)
        trunk.append(self.final_body)
        blocks = [self.body[g].body[b] for g in range(ng) for b in range(nb)]
        return dict(
            cfg=dict(cfg, style=self.style, no_group_conv=0, meta_relu=1, res_scale=1.0,
                     meta_hidden=(C_ // 2 if M <= 15 else (C_ - M) // 2 + M)),
            head=self.head[0], trunk=trunk, ups=[m for m in self.tail[0] if isinstance(m, nn.Conv2d)],
            tail=self.tail[1], ca=[blk.final_body.flat_params() for blk in blocks],
            ca_params=[blk.final_body.param_list() for blk in blocks],
            pa=[(blk.pa_node.pa[0], blk.pa_node.pa[2]) if blk.pa else None for blk in blocks],
            meta=[tuple(blk.q_node.fcs()) if blk.q_layer else None for blk in blocks])

    def packed(self, training=False):
        """Kernel-format parameters.  Rebuilt from scratch when a parameter's storage moved (`.to()`, new module);
        when only the values changed (optimizer.step(), load_state_dict) they are refreshed in place by one C call
        (`dfir_qrcan_repack`) driven by device pointer tables."""
        params = self.__dict__.get("_plist")
        if params is None:  # walking the module tree costs milliseconds per call: cached until `_apply` / load_state_dict
            params = self.__dict__["_plist"] = list(self.parameters())
        skey = (tuple(p.data_ptr() for p in params), self.precision, self.chunk_images, self.schedule)
        vers = tuple(p._version for p in params)
        pk = self._packed
        if pk is None or pk.key != skey or (pk.versions != vers and not pk.can_repack):
            if pk is not None:
                pk.close()
            pk = self._packed = PackedQrcan(self, skey)
            pk.versions = vers
        if training and not pk.train_ready:
            pk.enable_training(self)
            pk.versions = None
        if training and getattr(self, "cuda_graphs", True):
            pk.versions = None  # the step graph re-packs the parameters itself (deepfir_b200/train.py)
        elif pk.versions != vers:
            pk.repack()
            pk.versions = vers
        return pk

    def _apply(self, fn, *a, **k):
        self._packed = None
        self.__dict__.pop("_plist", None)
        return super()._apply(fn, *a, **k)

    def invalidate_packed(self):
        """Forces the kernel-format copies of the parameters (packed bf16 tiles, attention blobs, the wide-net planes) to
        be rebuilt from the nn.Parameters on the next forward.  The cache follows `Parameter._version`, which in-place
        writes through `.data` or any other alias do NOT bump (`p.data.copy_()`, weight averaging / EMA code, kernels that
        write the storage directly): call this after such a write.  `load_state_dict`, `.to()` and a train() -> eval()
        transition call it themselves."""
        pk = self.__dict__.get("_packed")
        if pk is None:
            pk = getattr(self, "_packed", None)
        if pk is not None:
            pk.versions = None
        self.__dict__.pop("_wide_key", None)

    def train(self, mode=True):
        if self.training and not mode:
            self.invalidate_packed()  # whatever trained the parameters may have written them behind autograd's back
        return super().train(mode)

    def load_state_dict(self, *a, **k):
        self.__dict__.pop("_plist", None)  # (assign=True replaces the Parameter objects)
        out = super().load_state_dict(*a, **k)
        self.invalidate_packed()
        return out


class QEDSRBlockParams(nn.Module):
    """ParamResBlock (architectures.py:332-356): conv-ReLU-conv, * res_scale, * meta-attention, += x."""

    def __init__(self, n_feats, n_params, q_layer_nonlinearity):
        super().__init__()
        self.body = nn.Sequential(_conv3(n_feats, n_feats), nn.ReLU(True), _conv3(n_feats, n_feats))
        self.attention_layer = MetaAttentionParams(n_feats, n_params, nonlinearity=q_layer_nonlinearity)


class QEDSR(QRCAN):
    """Q-EDSR on the B200 path (reference architectures.py:359-399): same constructor, parameter names, shapes
    and order.  Runs through the same C entry point as Q-RCAN with style NONE: one flat chain of blocks whose
    scale is res_scale * sigmoid(FC(FC(meta)))."""

    def __init__(self, in_features=3, out_features=3, num_features=64, input_para=1, num_blocks=16, scale=4,
                 res_scale=0.1, q_layer_nonlinearity=False, precision='bf16', chunk_images=0, schedule='linear',
                 **kwargs):
        nn.Module.__init__(self)
        if precision not in PRECISIONS:
            raise RuntimeError("precision must be 'bf16' or 'fp32'")
        bf16_planes = precision == 'bf16' and num_features % 64 == 0 and num_features <= 256
        if num_features % 8 or num_features > 256 or (256 % num_features and not bf16_planes):
            # the CUDA-core kernels (fp32 parity mode, training of wide nets) split 256-thread blocks by channel
            raise RuntimeError("Q-EDSR feature width %d is not supported in %s mode: fp32 mode and training need a divisor "
                               "of 256 that is a multiple of 8; bf16 inference also takes 192 (64-channel planes on the "
                               "tensor-core kernels)" % (num_features, precision))
        if schedule not in SCHEDULES:
            raise RuntimeError("schedule must be one of %s" % sorted(SCHEDULES))
        self.style = "none"
        self.scale = scale
        self.precision = precision
        self.chunk_images = chunk_images
        self.schedule = schedule
        M = input_para
        self.cfg = dict(n_resblocks=num_blocks, n_resgroups=1, n_feats=num_features, in_feats=in_features,
                        out_feats=out_features, scale=scale, reduction=16, num_metadata=M, style="none",
                        no_group_conv=1, meta_relu=int(bool(q_layer_nonlinearity)), res_scale=float(res_scale),
                        meta_hidden=(num_features // 2 if M <= 15 else (num_features - M) // 2 + M))
        self.head = _conv3(in_features, num_features)
        blocks = [QEDSRBlockParams(num_features, M, q_layer_nonlinearity) for _ in range(num_blocks)]
        self.final_body = _conv3(num_features, num_features)
        tail = [UpsamplerParams(scale, num_features), _conv3(num_features, out_features)]
        self.body = nn.Sequential(*blocks)
        self.tail = nn.Sequential(*tail)
        self._packed = None

    def forward(self, x, metadata):
        wide = self.precision == "bf16" and self.cfg["n_feats"] > 64 and self.cfg["n_feats"] % 64 == 0
        if not wide:
            return super().forward(x, metadata)
        # 128 / 256 features: 64-channel planes through the tensor-core kernels (deepfir_b200/wide.py)
        if not x.is_cuda:
            raise RuntimeError("deepfir_b200.QEDSR runs on a CUDA (sm_100a) device only: there is no CPU path")
        if torch.is_grad_enabled() and self.head.weight.requires_grad:
            # the tensor-core training kernels are specialised for 64 features: a training step of a wide net runs the
            # library's fp32 CUDA-core kernels (same schedule, divisors of 256); inference stays on the tensor cores
            if 256 % self.cfg["n_feats"]:
                raise NotImplementedError("training a %d-feature Q-EDSR is not supported (the fp32 training kernels need a "
                                          "divisor of 256): use 64, 128 or 256 features" % self.cfg["n_feats"])
            self.precision = "fp32"
            try:
                return super().forward(x, metadata)
            finally:
                self.precision = "bf16"
        from .wide import WideQEDSR
        params = self.__dict__.get("_plist")
        if params is None:
            params = self.__dict__["_plist"] = list(self.parameters())
        key = (tuple(p.data_ptr() for p in params), tuple(p._version for p in params))
        if self.__dict__.get("_wide_key") != key:
            self.__dict__["_wide"] = WideQEDSR(self)
            self.__dict__["_wide_key"] = key
        b = x.shape[0]
        attr = metadata.reshape(b, -1).to(device=x.device, dtype=torch.float32).contiguous()
        if attr.shape[1] != self.cfg["num_metadata"]:
            raise RuntimeError("metadata has %d entries per image, network expects %d" % (attr.shape[1], self.cfg["num_metadata"]))
        with torch.no_grad(), torch.cuda.device(x.device):
            return self.__dict__["_wide"].forward(x.to(torch.float32).contiguous(), attr)

    def _pack_spec(self):
        trunk = []
        for blk in self.body:
            trunk += [blk.body[0], blk.body[2]]
        trunk.append(self.final_body)
        return dict(cfg=self.cfg, head=self.head, trunk=trunk,
                    ups=[m for m in self.tail[0] if isinstance(m, nn.Conv2d)], tail=self.tail[1],
                    ca=[None for _ in self.body], ca_params=[None for _ in self.body],
                    meta=[tuple(blk.attention_layer.fcs()) for blk in self.body])


_HANDLES = {}
_NEXT = [1]


class PackedQrcan:
    """Kernel-format copy of a QRCAN's parameters + the `dfir_qrcan_net` descriptor."""

    def __init__(self, net, key):
        lib = _lib.load_library()
        self.key = key
        spec = net._pack_spec()
        cfg = spec["cfg"]
        dev = spec["head"].weight.device
        if dev.type != "cuda":
            raise RuntimeError("network parameters must live on a CUDA device")
        C_ = cfg["n_feats"]
        ng, nb = cfg["n_resgroups"], cfg["n_resblocks"]
        want_tc = net.precision == "bf16"
        if want_tc and C_ != 64:
            raise RuntimeError("the tensor-core path is specialised for 64 feature channels (use precision='fp32')")
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        with torch.no_grad(), torch.cuda.device(dev):
            trunk = spec["trunk"]
            ups = spec["ups"]
            r = 3 if net.scale == 3 else 2
            n_trunk = len(trunk)
            n_conv = n_trunk + len(ups) * r * r
            f32 = dict(device=dev, dtype=torch.float32)
            self.keep = []  # tensors whose storage the descriptor points into

            def keep(t):
                self.keep.append(t)
                return t

            conv_b = keep(torch.zeros(n_conv, C_, **f32))
            for i, m in enumerate(trunk):
                conv_b[i] = m.bias
            for t, m in enumerate(ups):
                conv_b[n_trunk + t * r * r: n_trunk + (t + 1) * r * r] = m.bias.reshape(C_, r * r).t()
            tail = spec["tail"]
            tail_b = keep(torch.zeros(16, **f32))
            tail_b[: tail.bias.numel()] = tail.bias
            head = spec["head"]
            head_w = keep(torch.empty(9 * head.in_channels * C_, **f32))
            _lib.check(lib.dfir_pack_conv3x3_f32(head.weight.contiguous().data_ptr(), head_w.data_ptr(), C_,
                                                 head.in_channels, stream), "pack head")
            head_b = keep(head.bias.detach().clone().contiguous())

            conv_w_bf16 = tail_w_bf16 = conv_w_f32 = up_w_f32 = tail_w_f32 = up_b = None
            if want_tc:
                wbytes = 9 * 64 * 128
                conv_w_bf16 = keep(torch.empty(n_conv * wbytes, device=dev, dtype=torch.uint8))
                for i, m in enumerate(trunk):
                    _lib.check(lib.dfir_pack_conv3x3_bf16(m.weight.contiguous().data_ptr(),
                                                          conv_w_bf16.data_ptr() + i * wbytes, 64, 64, 64, 0, 1,
                                                          stream), "pack trunk")
                for t, m in enumerate(ups):
                    w = m.weight.contiguous()
                    for s_ in range(r * r):
                        i = n_trunk + t * r * r + s_
                        _lib.check(lib.dfir_pack_conv3x3_bf16(w.data_ptr(), conv_w_bf16.data_ptr() + i * wbytes,
                                                              w.shape[0], 64, 64, s_, r * r, stream), "pack up")
                tail_w_bf16 = keep(torch.empty(9 * 16 * 128, device=dev, dtype=torch.uint8))
                _lib.check(lib.dfir_pack_conv3x3_bf16(tail.weight.contiguous().data_ptr(), tail_w_bf16.data_ptr(),
                                                      tail.weight.shape[0], 64, 16, 0, 1, stream), "pack tail")
            else:
                wsz = 9 * C_ * C_
                conv_w_f32 = keep(torch.empty(n_trunk * wsz, **f32))
                for i, m in enumerate(trunk):
                    _lib.check(lib.dfir_pack_conv3x3_f32(m.weight.contiguous().data_ptr(),
                                                         conv_w_f32.data_ptr() + i * wsz * 4, C_, C_, stream),
                               "pack trunk f32")
                usz = 9 * C_ * r * r * C_
                up_w_f32 = keep(torch.empty(max(1, len(ups)) * usz, **f32))
                for t, m in enumerate(ups):
                    _lib.check(lib.dfir_pack_conv3x3_f32(m.weight.contiguous().data_ptr(),
                                                         up_w_f32.data_ptr() + t * usz * 4, r * r * C_, C_, stream),
                               "pack up f32")
                up_b = keep(torch.cat([m.bias.reshape(-1) for m in ups]).contiguous())
                tail_w_f32 = keep(torch.empty(9 * C_ * tail.weight.shape[0], **f32))
                _lib.check(lib.dfir_pack_conv3x3_f32(tail.weight.contiguous().data_ptr(), tail_w_f32.data_ptr(),
                                                     tail.weight.shape[0], C_, stream), "pack tail f32")

            # attention blobs
            ca_rows = spec["ca"]          # per block: list of flat fp32 tensors, or None (no channel attention)
            metas = spec["meta"]          # per block: (fc1, fc2) or None
            nblk = len(ca_rows)
            if nblk and ca_rows[0] is not None:
                ca_blob = keep(torch.stack([torch.cat(rw) for rw in ca_rows]).to(**f32).contiguous())
                ca_stride = ca_blob.shape[1]
            else:
                ca_blob, ca_stride = keep(torch.zeros(8, **f32)), 0
            q_flags = [0 if m is None else 1 for m in metas]
            any_q = int(any(q_flags))
            # blocks scaled by a constant (EDSR's ResBlock: res_scale, no meta-attention): the scale buffer of a block
            # without a q layer holds out_scale, so an all-disabled table gives exactly that
            const_scale = bool(cfg.get("constant_block_scale")) and not any_q
            M = cfg["num_metadata"]
            hid = cfg["meta_hidden"]
            if any_q or const_scale:
                z = lambda *shape: torch.zeros(*shape, **f32)
                w1, b1, w2, b2 = z(nblk, hid, M), z(nblk, hid), z(nblk, C_, hid), z(nblk, C_)
                for i, m in enumerate(metas):
                    if m is not None:
                        f1, f2 = m
                        w1[i], b1[i] = f1.weight.reshape(hid, M), f1.bias
                        w2[i], b2[i] = f2.weight.reshape(C_, hid), f2.bias
                self.meta = [keep(t) for t in (w1, b1, w2, b2)]
                q_enabled = keep(torch.tensor(q_flags, device=dev, dtype=torch.int32))
                any_q = 1
            else:
                self.meta = [None] * 4
                q_enabled = None

            # pixel attention (PALayer) parameters, one row per block
            pas = spec.get("pa") or []
            pa_blob = None
            if any(p is not None for p in pas):
                if C_ != 64 or not all(p is not None for p in pas):
                    raise NotImplementedError("pixel attention is implemented for 64-feature Q-RCAN blocks")
                pa_blob = keep(torch.stack([torch.cat([f1.weight.reshape(-1), f1.bias.reshape(-1), f2.weight.reshape(-1),
                                                       f2.bias.reshape(-1)]) for f1, f2 in pas]).to(**f32).contiguous())

        style = STYLES[cfg["style"]]
        self.attr_size = C_ if cfg["style"] == "modulate" else M
        if any_q and self.attr_size != M:
            raise RuntimeError("style='modulate' cannot be combined with q layers (attribute size mismatch)")
        ptr = lambda t: (t.data_ptr() if t is not None else None)
        d = QrcanNet()
        d.n_groups, d.n_blocks, d.n_feats = ng, nb, C_
        d.scale, d.style, d.reduced = net.scale, style, max(1, C_ // cfg["reduction"])
        d.num_metadata, d.attr_size, d.meta_hidden = M, self.attr_size, hid
        d.in_feats, d.out_feats = cfg["in_feats"], cfg["out_feats"]
        d.q_enabled, d.any_q, d.chunk_images = ptr(q_enabled), any_q, int(net.chunk_images)
        d.schedule = SCHEDULES[net.schedule]
        d.no_group_conv, d.meta_relu, d.res_scale = int(cfg["no_group_conv"]), int(cfg["meta_relu"]), float(cfg["res_scale"])
        d.conv_w_bf16, d.tail_w_bf16 = ptr(conv_w_bf16), ptr(tail_w_bf16)
        d.conv_w_f32, d.up_w_f32, d.tail_w_f32, d.head_w_f32 = ptr(conv_w_f32), ptr(up_w_f32), ptr(tail_w_f32), ptr(head_w)
        d.conv_b, d.up_b, d.tail_b, d.head_b = ptr(conv_b), ptr(up_b), ptr(tail_b), ptr(head_b)
        d.ca_blob, d.ca_stride = ptr(ca_blob), int(ca_stride)
        d.meta_w1, d.meta_b1, d.meta_w2, d.meta_b2 = [ptr(t) for t in self.meta]
        d.pa_blob, d.pa_stride = ptr(pa_blob), (int(pa_blob.shape[1]) if pa_blob is not None else 0)
        self.desc = d
        self.device = dev
        self.scale = net.scale
        self.out_feats = cfg["out_feats"]
        self.precision = PRECISIONS[net.precision]
        self._ws = {}
        self.versions = None
        self.train_ready = False
        self.step_graphs = {}
        self.ws_owner = None
        # device pointer tables of the fp32 parameters: lets the C side refresh every kernel-format buffer in a few
        # launches (styles whose attention block has the 4-tensor layout; the others are rebuilt from Python)
        self.can_repack = "ca_params" in spec
        if self.can_repack:
            self._spec_params = dict(
                conv_w=[m.weight for m in trunk], conv_b=[m.bias for m in trunk],
                up_w=[m.weight for m in ups], up_b=[m.bias for m in ups],
                tail_w=tail.weight, tail_b=tail.bias, head_w=head.weight, head_b=head.bias,
                ca=(None if ca_stride == 0 else
                    [p for blk in spec["ca_params"] for p in (list(blk) + [None] * (8 - len(blk)))]),
                meta=(None if not any(q_flags) else
                      [t for m in metas for t in ((None,) * 4 if m is None else (m[0].weight, m[0].bias, m[1].weight, m[1].bias))]),
                pa=(None if pa_blob is None else
                    [t for f1, f2 in pas for t in (f1.weight, f1.bias, f2.weight, f2.bias)]))
            self.param_tables, self.params_struct = self._make_tables(lambda p: p.data_ptr())
        self.handle = _NEXT[0]
        _NEXT[0] += 1
        _HANDLES[self.handle] = self

    # ------------------------------------------------------------------ pointer tables / training
    def _make_tables(self, addr):
        """QrcanParams whose table entries are addr(parameter) (0 for absent tensors); returns (tensors kept alive,
        ctypes struct)"""
        sp = self._spec_params
        keep = {}

        def table(lst):
            if lst is None or len(lst) == 0:
                return None
            t = torch.tensor([0 if p is None else addr(p) for p in lst], dtype=torch.int64, device=self.device)
            return t

        ps = _lib.QrcanParams()
        for name in ("conv_w", "conv_b", "up_w", "up_b", "ca", "meta", "pa"):
            keep[name] = table(sp[name])
            setattr(ps, name, None if keep[name] is None else keep[name].data_ptr())
        for name in ("tail_w", "tail_b", "head_w", "head_b"):
            setattr(ps, name, addr(sp[name]))
        return keep, ps

    def repack(self):
        lib = _lib.load_library()
        with torch.cuda.device(self.device):
            stream = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            _lib.check(lib.dfir_qrcan_repack(C.byref(self.desc), C.byref(self.params_struct), self.precision,
                                             int(self.train_ready), stream), "repack")

    def enable_training(self, net):
        """Buffers that only a training step needs: data-gradient weight layouts, one flat gradient buffer with a
        view per parameter, and the gradient pointer tables."""
        if not self.can_repack:
            raise NotImplementedError("this network has no training path on the B200 library")
        d = self.desc
        dev = self.device
        C_ = d.n_feats
        r = 3 if self.scale == 3 else 2
        n_up = len(self._spec_params["up_w"])
        n_trunk = len(self._spec_params["conv_w"])
        with torch.cuda.device(dev):
            if self.precision == 0:
                self.wT = [torch.empty((n_trunk + n_up * r * r) * 9 * 64 * 128, device=dev, dtype=torch.uint8)]
                d.conv_wT_bf16 = self.wT[0].data_ptr()
            else:
                self.wT = [torch.empty(n_trunk * 9 * C_ * C_, device=dev, dtype=torch.float32),
                           torch.empty(max(1, n_up) * 9 * r * r * C_ * C_, device=dev, dtype=torch.float32)]
                d.conv_wT_f32, d.up_wT_f32 = self.wT[0].data_ptr(), self.wT[1].data_ptr()
            self.wT.append(torch.empty(9 * self.out_feats * C_, device=dev, dtype=torch.float32))
            d.tail_wT_f32 = self.wT[-1].data_ptr()
            params = list(net.parameters())
            offs, total = {}, 0
            for p in params:
                offs[id(p)] = total
                total += (p.numel() + 3) // 4 * 4  # 16-byte aligned views
            self.grad_params = params
            self.grad_flat = [torch.zeros(total, device=dev, dtype=torch.float32) for _ in range(2)]
            self.grad_views, self.grad_tables = [], []
            for flat in self.grad_flat:
                base = flat.data_ptr()
                self.grad_views.append([flat[offs[id(p)]: offs[id(p)] + p.numel()].view(p.shape) for p in params])
                self.grad_tables.append(self._make_tables(lambda p, base=base: base + 4 * offs[id(p)]))
        self.train_ready = True

    def train_workspace(self, B, H, W):
        k = ("train", B, H, W)
        ws = self._ws.get(k)
        if ws is None:
            lib = _lib.load_library()
            n = lib.dfir_qrcan_train_workspace_bytes(C.byref(self.desc), B, H, W, self.precision)
            if n == 0:
                raise NotImplementedError("this network configuration has no training path on the B200 library")
            self._ws.clear()
            ws = torch.empty(int(n), device=self.device, dtype=torch.uint8)
            self._ws[k] = ws
        return ws

    def workspace(self, B, H, W, precision):
        k = (B, H, W, precision)
        ws = self._ws.get(k)
        if ws is None:
            lib = _lib.load_library()
            n = lib.dfir_qrcan_workspace_bytes(C.byref(self.desc), B, H, W, precision)
            self._ws.clear()  # one live workspace per network keeps HBM use bounded
            ws = torch.empty(int(n), device=self.device, dtype=torch.uint8)
            self._ws[k] = ws
        return ws

    def close(self):
        _HANDLES.pop(self.handle, None)


def packed_from_handle(h):
    return _HANDLES[h]
