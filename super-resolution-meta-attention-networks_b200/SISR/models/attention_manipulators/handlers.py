"""Handlers of the metadata-conditioned networks (reference:
Code/SISR/models/attention_manipulators/handlers.py).  Same class names (the registry keys are derived from
them), constructor arguments and attributes; the networks they own are the B200 implementations in
`deepfir_b200`.  Extra optional `internal_params` keys understood here: `precision` ('bf16' | 'fp32') and
`chunk_images`; reference configs that do not carry them load unchanged."""
import time

import numpy as np
import torch
from torch import nn

from SISR.models.attention_manipulators import QModel
from deepfir_b200.han_san import QHAN, QSAN
from deepfir_b200.qrcan import QEDSR, QRCAN


class QRCANHandler(QModel):
    """Meta-attention RCAN: 10 residual groups x 20 RCABs by default.  `include_q_layer`,
    `selective_meta_blocks` (one bool per group) and `num_q_layers_inner_residual` place the
    meta-attention layers exactly as in the reference (handlers.py:7-40)."""

    def __init__(self, device, model_save_dir, eval_mode=False, lr=1e-4, scale=4, in_features=3, scheduler=None,
                 scheduler_params=None, style='modulate', perceptual=None, clamp=False, min_mu=-0.2,
                 max_mu=0.8, n_feats=64, **kwargs):
        super(QRCANHandler, self).__init__(device=device, model_save_dir=model_save_dir, eval_mode=eval_mode,
                                           **kwargs)
        self.net = QRCAN(scale=scale, in_feats=in_features, num_metadata=self.num_metadata,
                         n_feats=n_feats, style=style, **kwargs)
        self.colorspace = 'augmented_rgb'
        self.im_input = 'unmodified'
        self.activate_device()
        self.training_setup(lr, scheduler, scheduler_params, perceptual, device)
        self.model_name = 'qrcan'
        self.min_mu = min_mu
        self.max_mu = max_mu
        self.base_scaler = np.linspace(0, 1, n_feats)
        self.clamp = clamp
        self.style = style

    @staticmethod
    def gaussian(x, mu, sig=0.2):
        g = (1 / (np.sqrt(2 * np.pi) * sig)) * np.exp(-np.power(x - mu, 2.) / (2 * np.power(sig, 2.)))
        return torch.from_numpy(g).type(torch.float32)

    def scale_qpi(self, qpi):
        """'modulate' style: each image's scalar becomes an n_feats-bin Gaussian bump (ref :42-54)."""
        mu = (qpi * (self.max_mu - self.min_mu)) + self.min_mu
        rows = torch.stack([self.gaussian(self.base_scaler, mu[i].squeeze().numpy()) for i in range(mu.size(0))])
        if self.clamp:
            rows = torch.clamp(rows, 0, 1)
        return rows.unsqueeze(2).unsqueeze(3)


class QEDSRHandler(QModel):
    """Meta-attention EDSR (ref :57-76): ParamResBlock chain, every block scaled by its meta-attention vector."""

    def __init__(self, device, model_save_dir, eval_mode=False, lr=1e-4, scale=4, in_features=3, num_blocks=16,
                 num_features=64, res_scale=0.1, scheduler=None, scheduler_params=None, perceptual=None, **kwargs):
        super(QEDSRHandler, self).__init__(device=device, model_save_dir=model_save_dir, eval_mode=eval_mode,
                                           **kwargs)
        if num_features % 64 != 0:  # the tensor-core kernels work on 64-channel planes; other widths run fp32
            kwargs.setdefault('precision', 'fp32')
        self.net = QEDSR(scale=scale, in_features=in_features, num_features=num_features, num_blocks=num_blocks,
                         res_scale=res_scale, input_para=self.num_metadata, **kwargs)
        self.colorspace = 'augmented_rgb'
        self.im_input = 'unmodified'
        self.activate_device()
        self.model_name = 'qedsr'
        self.criterion = nn.L1Loss()
        self.training_setup(lr, scheduler, scheduler_params, perceptual, device)


class QSANHandler(QModel):
    """Meta-attention SAN (ref :79-153).  Evaluation always goes through `forward_chop`: four overlapping
    quadrants (10 px of overlap), each run through the network if smaller than `max_combined_im_size`, else
    chopped again; the inner halves are stitched back together."""

    def __init__(self, device, model_save_dir, eval_mode=False, lr=1e-4, scale=4, perceptual=None,
                 max_combined_im_size=160000, scheduler=None, scheduler_params=None, **kwargs):
        super(QSANHandler, self).__init__(device=device, model_save_dir=model_save_dir, eval_mode=eval_mode,
                                          **kwargs)
        extra = {k: kwargs[k] for k in ('precision', 'schedule') if k in kwargs}
        self.net = QSAN(scale=scale, input_para=self.num_metadata, **extra)  # like the reference: fixed 20 x 10 trunk
        self.scale = scale
        self.colorspace = 'rgb'
        self.im_input = 'unmodified'
        self.activate_device()
        self.training_setup(lr, scheduler, scheduler_params, perceptual, device)
        self.max_combined_im_size = max_combined_im_size
        self.model_name = 'qsan'

    def forward_chop(self, x, extra_channels, shave=10):
        b, c, h, w = x.size()
        h_half, w_half = h // 2, w // 2
        h_size, w_size = h_half + shave, w_half + shave
        quadrants = [x[:, :, 0:h_size, 0:w_size], x[:, :, 0:h_size, (w - w_size):w],
                     x[:, :, (h - h_size):h, 0:w_size], x[:, :, (h - h_size):h, (w - w_size):w]]
        if w_size * h_size < self.max_combined_im_size:
            sr = [self.run_chopped_eval(q, extra_channels) for q in quadrants]
        else:
            sr = [self.forward_chop(q, extra_channels, shave=shave) for q in quadrants]
        s = self.scale
        h, w, h_half, w_half, h_size, w_size = s * h, s * w, s * h_half, s * w_half, s * h_size, s * w_size
        output = x.new(b, c, h, w)
        output[:, :, 0:h_half, 0:w_half] = sr[0][:, :, 0:h_half, 0:w_half]
        output[:, :, 0:h_half, w_half:w] = sr[1][:, :, 0:h_half, (w_size - w + w_half):w_size]
        output[:, :, h_half:h, 0:w_half] = sr[2][:, :, (h_size - h + h_half):h_size, 0:w_half]
        output[:, :, h_half:h, w_half:w] = sr[3][:, :, (h_size - h + h_half):h_size, (w_size - w + w_half):w_size]
        return output

    def run_eval(self, x, y=None, request_loss=False, metadata=None, metadata_keys=None, timing=False, *args, **kwargs):
        extra_channels = self.generate_channels(x, metadata, metadata_keys).to(self.device)
        tic = time.perf_counter()
        sr_image = self.forward_chop(x, extra_channels)
        elapsed = time.perf_counter() - tic
        loss = self.criterion(sr_image, y) if request_loss else None
        return sr_image, loss, elapsed if timing else None

    def run_chopped_eval(self, x, extra_channels):
        return super().run_eval(x.contiguous(), y=None, request_loss=False, extra_channels=extra_channels)[0]


class QHANHandler(QModel):
    """Meta-attention HAN (ref :156-171): Q-RCAN groups + layer attention + channel-spatial attention."""

    def __init__(self, device, model_save_dir, eval_mode=False, lr=1e-4, scale=4, perceptual=None,
                 scheduler=None, scheduler_params=None, **kwargs):
        super(QHANHandler, self).__init__(device=device, model_save_dir=model_save_dir, eval_mode=eval_mode,
                                          **kwargs)
        extra = {k: kwargs[k] for k in ('precision', 'schedule') if k in kwargs}
        self.net = QHAN(scale=scale, num_metadata=self.num_metadata, **extra)
        self.colorspace = 'rgb'
        self.im_input = 'unmodified'
        self.activate_device()
        self.training_setup(lr, scheduler, scheduler_params, perceptual, device)
        self.model_name = 'qhan'
