"""Handlers of the non-meta baselines (reference: Code/SISR/models/advanced/handlers.py:7-39).  Same class names
(registry keys `edsr`, `rcan`), constructor arguments and attributes; the networks are the B200 implementations in
`deepfir_b200.baselines` — the Q-RCAN / Q-EDSR kernels with the meta-attention scale identically 1.  Optional extra
`internal_params`: `precision` ('bf16' | 'fp32')."""
from SISR.models import BaseModel
from deepfir_b200.baselines import EDSR, RCAN


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
