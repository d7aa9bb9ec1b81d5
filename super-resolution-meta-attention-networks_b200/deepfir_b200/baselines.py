"""The non-meta baselines RCAN and EDSR on the B200 path.

Reference: /root/reference/Code/SISR/models/advanced/architectures.py:13-243 and advanced/common.py:48-72.  Same
constructors, parameter names, shapes and registration order (checkpoints are interchangeable); the modules are
parameter containers and `forward` runs the Q-RCAN / Q-EDSR kernels of `libdfir_b200.so` with the meta-attention scale
identically 1: RCAB = conv-ReLU-conv + standard channel attention + skip, ResBlock = conv-ReLU-conv * res_scale + skip.
Inference and training (same autograd node as the Q-nets).
"""
import torch
from torch import nn

from .qrcan import PRECISIONS, SCHEDULES, QRCAN, UpsamplerParams, _conv3, _fc


class CALayerParams(nn.Module):
    """CALayer (advanced/architectures.py:13-33): avg-pool, FC-ReLU-FC-sigmoid, scale."""

    def __init__(self, channel, reduction=16):
        super().__init__()
        self.conv_du = nn.Sequential(_fc(channel, channel // reduction), nn.ReLU(inplace=True),
                                     _fc(channel // reduction, channel), nn.Sigmoid())

    def param_list(self):
        return [self.conv_du[0].weight, self.conv_du[0].bias, self.conv_du[2].weight, self.conv_du[2].bias]

    def flat_params(self):
        return [p.reshape(-1) for p in self.param_list()]


class RCABParams(nn.Module):
    def __init__(self, n_feat, reduction):
        super().__init__()
        self.body = nn.Sequential(_conv3(n_feat, n_feat), nn.ReLU(True), _conv3(n_feat, n_feat),
                                  CALayerParams(n_feat, reduction))


class ResidualGroupParams(nn.Module):
    def __init__(self, n_feat, reduction, n_resblocks):
        super().__init__()
        self.body = nn.Sequential(*([RCABParams(n_feat, reduction) for _ in range(n_resblocks)] + [_conv3(n_feat, n_feat)]))


class _NoMetaNet(QRCAN):
    """forward(x): the metadata argument of the Q-net entry point is a dummy (no layer reads it)"""

    def forward(self, x, metadata=None):
        if metadata is None:
            metadata = torch.zeros(x.shape[0], self.cfg["num_metadata"], 1, 1, device=x.device)
        return super().forward(x, metadata)


class RCAN(_NoMetaNet):
    def __init__(self, n_resblocks=20, n_resgroups=10, n_feats=64, in_feats=3, out_feats=3, scale=4, reduction=16,
                 res_scale=1.0, precision='bf16', chunk_images=0, schedule='linear', **kwargs):
        nn.Module.__init__(self)
        if precision not in PRECISIONS or schedule not in SCHEDULES:
            raise RuntimeError("unknown precision / schedule")
        self.style, self.scale, self.precision = "standard", scale, precision
        self.chunk_images, self.schedule = chunk_images, schedule
        self.cfg = dict(n_resblocks=n_resblocks, n_resgroups=n_resgroups, n_feats=n_feats, in_feats=in_feats,
                        out_feats=out_feats, scale=scale, reduction=reduction, num_metadata=1,
                        include_pixel_attention=False)
        head = [_conv3(in_feats, n_feats)]
        body = [ResidualGroupParams(n_feats, reduction, n_resblocks) for _ in range(n_resgroups)]
        body.append(_conv3(n_feats, n_feats))
        tail = [UpsamplerParams(scale, n_feats), _conv3(n_feats, out_feats)]
        self.head = nn.Sequential(*head)
        self.body = nn.Sequential(*body)
        self.tail = nn.Sequential(*tail)
        self._packed = None

    def _pack_spec(self):
        cfg = self.cfg
        ng, nb, C_ = cfg["n_resgroups"], cfg["n_resblocks"], cfg["n_feats"]
        trunk, blocks = [], []
        for g in range(ng):
            grp = self.body[g].body
            for b in range(nb):
                trunk += [grp[b].body[0], grp[b].body[2]]
                blocks.append(grp[b].body[3])
            trunk.append(grp[nb])
        trunk.append(self.body[ng])
        return dict(cfg=dict(cfg, style="standard", no_group_conv=0, meta_relu=1, res_scale=1.0, meta_hidden=C_ // 2),
                    head=self.head[0], trunk=trunk, ups=[m for m in self.tail[0] if isinstance(m, nn.Conv2d)],
                    tail=self.tail[1], ca=[blk.flat_params() for blk in blocks],
                    ca_params=[blk.param_list() for blk in blocks], pa=[None for _ in blocks],
                    meta=[None for _ in blocks])


class ResBlockParams(nn.Module):
    def __init__(self, n_feats):
        super().__init__()
        self.body = nn.Sequential(_conv3(n_feats, n_feats), nn.ReLU(True), _conv3(n_feats, n_feats))


class EDSR(_NoMetaNet):
    def __init__(self, in_features=3, out_features=3, net_features=64, num_blocks=16, scale=4, res_scale=0.1,
                 precision='bf16', chunk_images=0, schedule='linear', **kwargs):
        nn.Module.__init__(self)
        if precision not in PRECISIONS or schedule not in SCHEDULES:
            raise RuntimeError("unknown precision / schedule")
        self.style, self.scale, self.precision = "none", scale, precision
        self.chunk_images, self.schedule = chunk_images, schedule
        self.cfg = dict(n_resblocks=num_blocks, n_resgroups=1, n_feats=net_features, in_feats=in_features,
                        out_feats=out_features, scale=scale, reduction=16, num_metadata=1, style="none",
                        no_group_conv=1, meta_relu=0, res_scale=float(res_scale), meta_hidden=net_features // 2,
                        constant_block_scale=True)
        head = [_conv3(in_features, net_features)]
        body = [ResBlockParams(net_features) for _ in range(num_blocks)] + [_conv3(net_features, net_features)]
        tail = [UpsamplerParams(scale, net_features), _conv3(net_features, out_features)]
        self.head = nn.Sequential(*head)
        self.body = nn.Sequential(*body)
        self.tail = nn.Sequential(*tail)
        self._packed = None

    def _pack_spec(self):
        nb = self.cfg["n_resblocks"]
        trunk = []
        for b in range(nb):
            trunk += [self.body[b].body[0], self.body[b].body[2]]
        trunk.append(self.body[nb])
        return dict(cfg=self.cfg, head=self.head[0], trunk=trunk,
                    ups=[m for m in self.tail[0] if isinstance(m, nn.Conv2d)], tail=self.tail[1],
                    ca=[None] * nb, ca_params=[None] * nb, pa=[None] * nb, meta=[None] * nb)


# ====================================================================================================
# SAN / HAN (advanced/architectures.py:244-377): the Q-SAN / Q-HAN data flow without meta-attention layers
# ====================================================================================================
from .han_san import (CSAMParams, LAMParams, NonlocalCAParams, QHAN, QSAN, SOCAParams)  # noqa: E402


class RBParams(nn.Module):
    """RB (advanced/SAN_blocks.py:339-363): conv-ReLU-conv + x."""

    def __init__(self, n_feat):
        super().__init__()
        self.conv_first = nn.Sequential(_conv3(n_feat, n_feat), nn.ReLU(inplace=True), _conv3(n_feat, n_feat))


class LSRAGParams(nn.Module):
    """LSRAG (advanced/SAN_blocks.py:366-412): RBs, SOCA, conv, + group input (`gamma` is serialised but unused)."""

    def __init__(self, n_feat, reduction, n_resblocks):
        super().__init__()
        self.rcab = nn.ModuleList([RBParams(n_feat) for _ in range(n_resblocks)])
        self.soca = SOCAParams(n_feat, reduction=reduction)
        self.conv_last = _conv3(n_feat, n_feat)
        self.gamma = nn.Parameter(torch.zeros(1))


class SAN(QSAN):
    def __init__(self, n_resgroups=20, n_resblocks=10, n_feats=64, reduction=16, scale=4, rgb_range=255, n_colors=3,
                 res_scale=1, precision='bf16', schedule='linear', **kwargs):
        nn.Module.__init__(self)
        if precision not in PRECISIONS or schedule not in SCHEDULES:
            raise RuntimeError("unknown precision / schedule")
        self.precision, self.schedule, self.scale, self.style = precision, schedule, scale, "none"
        self.cfg = dict(n_resblocks=n_resblocks, n_resgroups=n_resgroups, n_feats=n_feats, in_feats=n_colors,
                        out_feats=n_colors, scale=scale, reduction=reduction, num_metadata=1, style="none",
                        no_group_conv=1, meta_relu=1, res_scale=1.0, meta_hidden=n_feats // 2,
                        constant_block_scale=True)
        head = [_conv3(n_colors, n_feats)]
        self.gamma = nn.Parameter(torch.zeros(1))
        self.RG = nn.ModuleList([LSRAGParams(n_feats, reduction, n_resblocks) for _ in range(n_resgroups)])
        self.conv_last = _conv3(n_feats, n_feats)
        tail = [UpsamplerParams(scale, n_feats), _conv3(n_feats, n_colors)]
        self.non_local = NonlocalCAParams(in_feat=n_feats, inter_feat=n_feats // 8, reduction=8)
        self.head = nn.Sequential(*head)
        self.tail = nn.Sequential(*tail)

    def forward(self, x, metadata=None):  # (QSAN._pack_spec already copes with blocks that own no q_layer)
        if metadata is None:
            metadata = torch.zeros(x.shape[0], 1, 1, 1, device=x.device)
        return super().forward(x, metadata)


class HAN(QHAN):
    def __init__(self, n_resgroups=10, n_resblocks=20, n_feats=64, reduction=16, scale=4, n_colors=3, res_scale=1.0,
                 precision='bf16', schedule='linear', **kwargs):
        nn.Module.__init__(self)
        if precision not in PRECISIONS or schedule not in SCHEDULES:
            raise RuntimeError("unknown precision / schedule")
        self.precision, self.schedule, self.scale, self.style = precision, schedule, scale, "standard"
        self.cfg = dict(n_resblocks=n_resblocks, n_resgroups=n_resgroups, n_feats=n_feats, in_feats=n_colors,
                        out_feats=n_colors, scale=scale, reduction=reduction, num_metadata=1, style="standard",
                        no_group_conv=0, meta_relu=1, res_scale=1.0, meta_hidden=n_feats // 2)
        head = [_conv3(n_colors, n_feats)]
        body = [ResidualGroupParams(n_feats, reduction, n_resblocks) for _ in range(n_resgroups)]
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
        trunk, blocks = [], []
        for g in range(ng):
            grp = self.body[g].body
            for b in range(nb):
                trunk += [grp[b].body[0], grp[b].body[2]]
                blocks.append(grp[b].body[3])
            trunk.append(grp[nb])
        trunk.append(self.body[ng])
        return dict(cfg=cfg, head=self.head[0], trunk=trunk,
                    ups=[m for m in self.tail[0] if isinstance(m, nn.Conv2d)], tail=self.tail[1],
                    ca=[blk.flat_params() for blk in blocks], ca_params=[blk.param_list() for blk in blocks],
                    meta=[None for _ in blocks])

    def forward(self, x, metadata=None):
        if metadata is None:
            metadata = torch.zeros(x.shape[0], 1, 1, 1, device=x.device)
        return super().forward(x, metadata)
