"""`QModel`: handler base for networks modulated by image metadata (reference:
Code/SISR/models/attention_manipulators/__init__.py:6-118).  Turns the DataLoader's metadata batch into the
(B, M, 1, 1) fp32 `attributes` tensor the meta-attention layers consume and threads it into `net.forward`."""
import numpy as np
import torch

from SISR.models import BaseModel


class QModel(BaseModel):
    def __init__(self, metadata=None, **kwargs):
        self.style = None             # only Q-RCAN uses a channel-attention style
        self.channel_concat = False   # Q-nets never concatenate metadata with the input image
        if metadata is None:
            self.metadata = ['qpi']
            self.num_metadata = 1
        else:
            self.metadata = metadata
            extra = 0
            if 'all' in metadata:
                extra += 39           # all CelebA attributes
            if 'blur_kernel' in metadata:
                extra += 9            # a 10-D PCA blur code counts as one key
            elif 'unmodified_blur_kernel' in metadata:
                extra += 440
            self.num_metadata = len(metadata) + extra
        super(QModel, self).__init__(**kwargs)

    # every Q-handler ends its constructor the same way (ref handlers.py :31-35, :69-73, :90-94, :166-170): record the
    # colour space / input convention, move the parameters to the device, create optimizer + scheduler, name the model
    model_key = None
    colour = 'rgb'

    def finish_setup(self, lr, scheduler, scheduler_params, perceptual, device):
        self.colorspace = self.colour
        self.im_input = 'unmodified'
        self.activate_device()
        self.training_setup(lr, scheduler, scheduler_params, perceptual, device)
        self.model_name = self.model_key

    def generate_channels(self, x, metadata, keys):
        """metadata (B, K) + keys (K tuples, DataLoader-collated) -> (B, M, 1, 1) fp32 (ref :30-51).
        Vectorised: one masked gather instead of the reference's per-image Python loop."""
        if metadata is None:
            raise RuntimeError('Metadata needs to be specified for this network to run properly.')
        md = torch.as_tensor(np.asarray(metadata) if not torch.is_tensor(metadata) else metadata)
        batch = x.size(0)
        if len(keys) == 1:
            picked = md.reshape(batch, -1)
        else:
            if 'all' in self.metadata:
                mask = torch.ones(self.num_metadata, dtype=torch.bool)
            else:
                mask = torch.tensor([key[0] in self.metadata for key in keys], dtype=torch.bool)
            picked = md[:, mask.to(md.device)]
        extra_channels = (torch.ones(batch, self.num_metadata) * picked.to('cpu', torch.float32))
        extra_channels = extra_channels.unsqueeze(2).unsqueeze(3)
        if self.style == 'modulate':
            extra_channels = self.scale_qpi(extra_channels)
        return extra_channels

    def channel_concat_logic(self, x, extra_channels, metadata, metadata_keys):
        if extra_channels is None:
            extra_channels = self.generate_channels(x, metadata, metadata_keys)
            if not self.channel_concat and self.device != extra_channels.device:
                extra_channels = extra_channels.to(self.device)
        input_data = torch.cat((x, extra_channels), 1) if self.channel_concat else x
        return input_data, extra_channels

    def run_train(self, x, y, metadata=None, extra_channels=None, metadata_keys=None, *args, **kwargs):
        input_data, extra_channels = self.channel_concat_logic(x, extra_channels, metadata, metadata_keys)
        return super().run_train(input_data, y, extra_channels=extra_channels, **kwargs)

    def run_eval(self, x, y=None, request_loss=False, metadata=None, metadata_keys=None,
                 extra_channels=None, *args, **kwargs):
        input_data, extra_channels = self.channel_concat_logic(x, extra_channels, metadata, metadata_keys)
        return super().run_eval(input_data, y, request_loss=request_loss, extra_channels=extra_channels, **kwargs)

    def run_forensic(self, x, metadata=None, metadata_keys=None, extra_channels=None, *args, **kwargs):
        input_data, extra_channels = self.channel_concat_logic(x, extra_channels, metadata, metadata_keys)
        return super().run_forensic(input_data, qpi=extra_channels)

    def run_model(self, x, extra_channels=None, *args, **kwargs):
        return self.net.forward(x, metadata=extra_channels)
