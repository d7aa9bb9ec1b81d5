"""Handlers of the metadata-conditioned networks on the B200 path.

API mirror of the reference's Code/SISR/models/attention_manipulators/handlers.py: the registry derives the model keys
(`qrcan`, `qedsr`, `qsan`, `qhan`) from these class names, and `ModelInterface` / the training and evaluation loops rely
on the constructor arguments and attributes kept here.  The networks the handlers own are the B200 implementations of
`deepfir_b200`.  Optional extra `internal_params` keys: `precision` ('bf16' | 'fp32'), `schedule`, `chunk_images`;
reference TOMLs that do not carry them load unchanged.
"""
import math
import time

import torch
from torch import nn

from SISR.models.attention_manipulators import QModel
from deepfir_b200.han_san import QHAN, QSAN
from deepfir_b200.qrcan import QEDSR, QRCAN

_B200_KEYS = ('precision', 'schedule')


class QRCANHandler(QModel):
    """Meta-attention RCAN (ref :7-54): 10 residual groups x 20 RCABs by default; `include_q_layer`,
    `selective_meta_blocks` (one flag per group) and `num_q_layers_inner_residual` place the meta-attention layers;
    `style` picks the QCALayer variant ('modulate' feeds a Gaussian bump built from the scalar metadata)."""

    model_key = 'qrcan'
    colour = 'augmented_rgb'

    def __init__(self, device, model_save_dir, eval_mode=False, lr=1e-4, scale=4, in_features=3, scheduler=None,
                 scheduler_params=None, style='modulate', perceptual=None, clamp=False, min_mu=-0.2,
                 max_mu=0.8, n_feats=64, **kwargs):
        super().__init__(device=device, model_save_dir=model_save_dir, eval_mode=eval_mode, **kwargs)
        self.net = QRCAN(scale=scale, in_feats=in_features, num_metadata=self.num_metadata, n_feats=n_feats, style=style,
                         **kwargs)
        self.finish_setup(lr, scheduler, scheduler_params, perceptual, device)
        self.style = style
        self.clamp = clamp
        self.min_mu, self.max_mu = min_mu, max_mu
        self.base_scaler = torch.linspace(0, 1, n_feats, dtype=torch.float64).numpy()

    @staticmethod
    def gaussian(x, mu, sig=0.2):
        """normal density with mean `mu`, evaluated on the bin centres `x` (ref :37-40)"""
        grid = torch.as_tensor(x, dtype=torch.float64)
        mu = torch.as_tensor(mu, dtype=torch.float64)
        dens = torch.exp(-(grid - mu) ** 2 / (2.0 * sig ** 2)) / (math.sqrt(2.0 * math.pi) * sig)
        return dens.to(torch.float32)

    def scale_qpi(self, qpi):
        """'modulate' style (ref :42-54): every image's scalar in [0,1] moves the centre of a Gaussian over the
        n_feats bins; returns (B, n_feats, 1, 1).  Vectorised over the batch."""
        centres = (qpi.reshape(qpi.size(0), 1) * (self.max_mu - self.min_mu) + self.min_mu).to(torch.float64)
        bumps = self.gaussian(torch.as_tensor(self.base_scaler).reshape(1, -1), centres)
        if self.clamp:
            bumps = bumps.clamp(0, 1)
        return bumps[:, :, None, None]


class QEDSRHandler(QModel):
    """Meta-attention EDSR (ref :57-76): a chain of ParamResBlocks, each scaled by its meta-attention vector."""

    model_key = 'qedsr'
    colour = 'augmented_rgb'

    def __init__(self, device, model_save_dir, eval_mode=False, lr=1e-4, scale=4, in_features=3, num_blocks=16,
                 num_features=64, res_scale=0.1, scheduler=None, scheduler_params=None, perceptual=None, **kwargs):
        super().__init__(device=device, model_save_dir=model_save_dir, eval_mode=eval_mode, **kwargs)
        if num_features % 64:  # the tensor-core kernels work on 64-channel planes; other widths take the fp32 kernels
            kwargs.setdefault('precision', 'fp32')
        self.net = QEDSR(scale=scale, in_features=in_features, num_features=num_features, num_blocks=num_blocks,
                         res_scale=res_scale, input_para=self.num_metadata, **kwargs)
        self.criterion = nn.L1Loss()
        self.finish_setup(lr, scheduler, scheduler_params, perceptual, device)


class QSANHandler(QModel):
    """Meta-attention SAN (ref :79-153).  Like the reference, evaluation never sees the whole image: it is cut into four
    overlapping quadrants (recursively while a quadrant has `max_combined_im_size` pixels or more), each quadrant runs
    through the network, and the non-overlapping parts are stitched together."""

    model_key = 'qsan'

    def __init__(self, device, model_save_dir, eval_mode=False, lr=1e-4, scale=4, perceptual=None,
                 max_combined_im_size=160000, scheduler=None, scheduler_params=None, **kwargs):
        super().__init__(device=device, model_save_dir=model_save_dir, eval_mode=eval_mode, **kwargs)
        self.net = QSAN(scale=scale, input_para=self.num_metadata, **{k: kwargs[k] for k in _B200_KEYS if k in kwargs})
        self.scale = scale
        self.max_combined_im_size = max_combined_im_size
        self.finish_setup(lr, scheduler, scheduler_params, perceptual, device)

    def forward_chop(self, x, extra_channels, shave=10):
        """Quadrant (i, j) covers rows [0, h/2 + shave) or [h - h/2 - shave, h) (same for columns); of its SR result the
        rows / columns nearest to its own image corner are kept: h/2 (or the remaining h - h/2) of them.
        The four quadrants have the same size, so they run as ONE batch of 4 x B images (the reference runs them one after
        the other, handlers.py:99-150 — same values, a quarter of the launches); x, the quadrants and the stitched result
        stay on the device."""
        batch, chans, height, width = x.shape
        half = (height // 2, width // 2)
        size = (half[0] + shave, half[1] + shave)
        full = (height, width)

        def span(axis, far):  # LR slice of a quadrant along one axis
            return slice(full[axis] - size[axis], full[axis]) if far else slice(0, size[axis])

        corners = [(0, 0), (0, 1), (1, 0), (1, 1)]
        quads = [x[:, :, span(0, fr), span(1, fc)] for fr, fc in corners]
        if size[0] * size[1] < self.max_combined_im_size:
            sr_all = self.run_chopped_eval(torch.cat(quads, dim=0),
                                           None if extra_channels is None else torch.cat([extra_channels] * 4, dim=0))
            srs = [sr_all[i * batch:(i + 1) * batch] for i in range(4)]
        else:
            srs = [QSANHandler.forward_chop(self, q, extra_channels, shave=shave) for q in quads]
        s = self.scale
        out = srs[0].new_empty(batch, srs[0].shape[1], s * height, s * width)
        for (far_r, far_c), sr in zip(corners, srs):
            dst, src = [], []
            for axis, far in ((0, far_r), (1, far_c)):
                cut, whole, tile = s * half[axis], s * full[axis], s * size[axis]
                dst.append(slice(cut, whole) if far else slice(0, cut))
                src.append(slice(tile - (whole - cut), tile) if far else slice(0, cut))
            out[:, :, dst[0], dst[1]] = sr[:, :, src[0], src[1]]
        return out

    def run_eval(self, x, y=None, request_loss=False, metadata=None, metadata_keys=None, timing=False, *args, **kwargs):
        attributes = self.generate_channels(x, metadata, metadata_keys).to(self.device)
        started = time.perf_counter()
        if kwargs.pop('shard_tiles', False):
            # one process per GPU (torchrun): the 4 * B quadrants are spread over the ranks, one all-reduce stitches them
            from deepfir_b200.sharding import env_rank_world, run_chopped_sharded
            rank, world, _ = env_rank_world()
            size = (x.shape[2] // 2 + 10) * (x.shape[3] // 2 + 10)
            if size >= self.max_combined_im_size:
                raise NotImplementedError("shard_tiles: quadrants of %d pixels need a second chop level (max_combined_im_size "
                                          "= %d); only one level is sharded" % (size, self.max_combined_im_size))
            sr_image = self._to_host(run_chopped_sharded(self.run_chopped_eval, x.to(self.device), attributes, self.scale,
                                                         rank, world, out_channels=x.shape[1]))
        else:
            sr_image = self._to_host(self.forward_chop(x.to(self.device), attributes))  # one H2D, one D2H
        elapsed = time.perf_counter() - started
        return sr_image, (self.criterion(sr_image, y) if request_loss else None), (elapsed if timing else None)

    def run_chopped_eval(self, x, extra_channels):
        return super().run_eval(x.contiguous(), y=None, request_loss=False, extra_channels=extra_channels,
                                keep_on_device=True)[0]


class QHANHandler(QModel):
    """Meta-attention HAN (ref :156-171): Q-RCAN groups + layer attention + channel-spatial attention."""

    model_key = 'qhan'

    def __init__(self, device, model_save_dir, eval_mode=False, lr=1e-4, scale=4, perceptual=None,
                 scheduler=None, scheduler_params=None, **kwargs):
        super().__init__(device=device, model_save_dir=model_save_dir, eval_mode=eval_mode, **kwargs)
        self.net = QHAN(scale=scale, num_metadata=self.num_metadata, **{k: kwargs[k] for k in _B200_KEYS if k in kwargs})
        self.finish_setup(lr, scheduler, scheduler_params, perceptual, device)
