"""Handlers of the non-meta baselines (reference: Code/SISR/models/advanced/handlers.py:7-39).  Same class names
(registry keys `edsr`, `rcan`, `han`, `san`), constructor arguments and attributes; the networks are the B200 implementations in
`deepfir_b200.baselines` — the Q-RCAN / Q-EDSR kernels with the meta-attention scale identically 1.  Optional extra
`internal_params`: `precision` ('bf16' | 'fp32')."""
import time

from SISR.models import BaseModel
from SISR.models.attention_manipulators.handlers import QSANHandler
from deepfir_b200.baselines import EDSR, HAN, RCAN, SAN


class EDSRHandler(BaseModel):
    """EDSR (ref :7-23): ResBlock chain, `res_scale` 0.1."""

    def __init__(self, device, model_save_dir, eval_mode=False, lr=1e-4, scale=4, in_features=3, hr_data_loc=None,
                 scheduler=None, scheduler_params=None, perceptual=None, num_features=64, num_blocks=16, res_scale=0.1,
                 **kwargs):
        extra = {k: kwargs.pop(k) for k in ('precision', 'schedule', 'chunk_images') if k in kwargs}
        super(EDSRHandler, self).__init__(device=device, model_save_dir=model_save_dir, eval_mode=eval_mode, **kwargs)
        if num_features != 64:  # the 64-channel-plane path of wide nets needs the meta MLPs' layout: fp32 kernels instead
            extra.setdefault('precision', 'fp32')
        self.net = EDSR(scale=scale, in_features=in_features, net_features=num_features, num_blocks=num_blocks,
                        res_scale=res_scale, **extra)
        self.colorspace = 'rgb'
        self.im_input = 'unmodified'
        self.activate_device()
        self.training_setup(lr, scheduler, scheduler_params, perceptual, device)
        self.model_name = 'edsr'


class RCANHandler(BaseModel):
    """RCAN (ref :26-39): 10 residual groups x 20 RCABs, architecture locked like the reference."""

    def __init__(self, device, model_save_dir, eval_mode=False, lr=1e-4, scale=4, in_features=3, perceptual=None,
                 scheduler=None, scheduler_params=None, **kwargs):
        extra = {k: kwargs.pop(k) for k in ('precision', 'schedule', 'chunk_images') if k in kwargs}
        super(RCANHandler, self).__init__(device=device, model_save_dir=model_save_dir, eval_mode=eval_mode, **kwargs)
        self.net = RCAN(scale=scale, in_feats=in_features, **extra)
        self.colorspace = 'rgb'
        self.im_input = 'unmodified'
        self.activate_device()
        self.training_setup(lr, scheduler, scheduler_params, perceptual, device)
        self.model_name = 'rcan'


class HANHandler(BaseModel):
    """HAN (ref :42-55): RCAN groups + layer attention + channel-spatial attention, architecture locked."""

    def __init__(self, device, model_save_dir, eval_mode=False, lr=1e-4, scale=4, perceptual=None, scheduler=None,
                 scheduler_params=None, **kwargs):
        extra = {k: kwargs.pop(k) for k in ('precision', 'schedule') if k in kwargs}
        super(HANHandler, self).__init__(device=device, model_save_dir=model_save_dir, eval_mode=eval_mode, **kwargs)
        self.net = HAN(scale=scale, **extra)
        self.colorspace = 'rgb'
        self.im_input = 'unmodified'
        self.activate_device()
        self.training_setup(lr, scheduler, scheduler_params, perceptual, device)
        self.model_name = 'han'


class SANHandler(BaseModel):
    """SAN (ref :58-129).  Evaluation cuts the image into four overlapping quadrants (recursively while a quadrant has
    `max_combined_im_size` pixels or more) and stitches the results, exactly like the Q-SAN handler."""

    def __init__(self, device, model_save_dir, eval_mode=False, lr=1e-4, scale=4, perceptual=None,
                 max_combined_im_size=160000, scheduler=None, scheduler_params=None, **kwargs):
        extra = {k: kwargs.pop(k) for k in ('precision', 'schedule') if k in kwargs}
        super(SANHandler, self).__init__(device=device, model_save_dir=model_save_dir, eval_mode=eval_mode, **kwargs)
        self.net = SAN(scale=scale, **extra)
        self.scale = scale
        self.colorspace = 'rgb'
        self.im_input = 'unmodified'
        self.activate_device()
        self.training_setup(lr, scheduler, scheduler_params, perceptual, device)
        self.max_combined_im_size = max_combined_im_size
        self.model_name = 'san'

    def forward_chop(self, x, shave=10, **kw):
        # shares the quadrant / stitch logic of the Q-SAN handler (which recurses with an `extra_channels` slot)
        if shave is None or not isinstance(shave, int):
            shave = kw.get('shave', 10)
        return QSANHandler.forward_chop(self, x, None, shave=shave)

    def run_chopped_eval(self, x, extra_channels=None):
        return BaseModel.run_eval(self, x.contiguous(), request_loss=False, keep_on_device=True)[0]

    def run_eval(self, x, y=None, request_loss=False, metadata=None, metadata_keys=None, timing=False, *args, **kwargs):
        started = time.perf_counter()
        sr_image = self._to_host(self.forward_chop(x.to(self.device)))  # four quadrants as one batch, one H2D, one D2H
        elapsed = time.perf_counter() - started
        return sr_image, (self.criterion(sr_image, y) if request_loss else None), (elapsed if timing else None)
