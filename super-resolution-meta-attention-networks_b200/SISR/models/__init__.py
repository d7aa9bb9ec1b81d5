"""Model registry, client interface and handler base class.

Behavioural mirror of the reference's ``Code/SISR/models/__init__.py``:
  * registry (``available_models``): every ``class <Name>Handler`` found by ``ast`` in
    ``SISR/models/<dir>/handlers.py`` is registered under ``<name>.lower()``           (ref :20-30)
  * ``ModelInterface``: experiment folders, checkpoint selection, colour-space post-processing (ref :33-254)
  * ``BaseModel``: optimizer / scheduler / checkpoint dict layout / run_train / run_eval     (ref :257-574)

The checkpoint file format is unchanged: ``torch.save`` of a dict with the keys ``network`` (the net's
``state_dict``: same names, OIHW fp32 shapes), ``optimizer``, ``model_name``, ``model_epoch`` and
``scheduler_G``; files are ``<experiment>/saved_models/train_model_<epoch>``.
"""
import ast
import glob
import math
import os
import time
from collections import OrderedDict
from pydoc import locate

import numpy as np
import torch
import torch.nn as nn
import torch.optim as optim

from SISR.configuration.constants import base_directory
from sr_tools.helper_functions import create_dir_if_empty, read_metadata
from sr_tools.image_manipulation import ycbcr_convert

# ---------------------------------------------------------------------------------------------
# registry
# ---------------------------------------------------------------------------------------------
model_dir = os.path.join(base_directory, 'SISR', 'models')
available_models = {}


def _scan_handlers():
    for entry in sorted(os.scandir(model_dir), key=lambda e: e.name):
        if not entry.is_dir() or '__' in entry.name:
            continue
        with open(os.path.join(model_dir, entry.name, 'handlers.py'), 'r') as fh:
            tree = ast.parse(fh.read())
        for node in ast.walk(tree):
            if isinstance(node, ast.ClassDef):
                key = node.name.split('Handler')[0].lower()
                available_models[key] = 'SISR.models.%s.handlers.%s' % (entry.name, node.name)


_scan_handlers()


def _toml_load(path):
    try:
        import toml
        return toml.load(path)
    except ImportError:  # python >= 3.11
        import tomllib
        with open(path, 'rb') as fh:
            return tomllib.load(fh)


class ModelInterface:
    """Client-side interface: builds the handler named in the experiment's config, loads/saves
    checkpoints, and formats network outputs (ref :33-254; same constructor arguments)."""

    def __init__(self, model_loc, experiment, gpu='off', sp_gpu=0, mode='eval', new_params=None,
                 load_epoch=None, scale=None, save_subdir=None, new_branch=False):
        sub = (lambda d: os.path.join(d, save_subdir)) if save_subdir is not None else (lambda d: d)
        self.experiment = experiment
        self.base_folder = os.path.abspath(os.path.join(model_loc, experiment))
        self.logs = os.path.abspath(os.path.join(self.base_folder, sub('result_outputs')))
        self.saved_models = os.path.abspath(os.path.join(self.base_folder, sub('saved_models')))
        self.mode = mode
        load_override = os.path.dirname(self.saved_models) if new_branch else None

        if mode == 'train':
            create_dir_if_empty(self.base_folder, self.logs, self.saved_models)
            if new_params is None and load_epoch is None:
                raise RuntimeError('Need to specify model parameters to train a new model.')
        elif mode == 'eval' and load_epoch is None:
            raise RuntimeError('Need to specify which model epoch to load.')

        if load_epoch is None:
            self.model_epoch = 0
            self.metadata = new_params
        elif not glob.glob(os.path.join(self.base_folder, '*.toml')):
            self.metadata = self._legacy_model_setup(experiment, self.base_folder, scale)
        else:
            self.metadata = _toml_load(os.path.join(self.base_folder, 'config.toml'))['model']

        self.name = self.metadata['name']
        if self.name == 'qpircan':  # legacy alias
            self.name = 'qrcan'
        if scale is not None and scale != self.metadata['internal_params']['scale']:
            raise Exception('The model loaded has been trained for a different scale, '
                            'and cannot produce the requested images.')

        self.device = sp_gpu if (gpu != 'off' and torch.cuda.is_available()) else torch.device('cpu')
        self.model = self.define_model(name=self.name, model_save_dir=self.saved_models, device=self.device,
                                       eval_mode=(mode == 'eval'), **self.metadata['internal_params'])

        if load_epoch is not None:
            if load_epoch in ('best', 'last'):
                import pandas as pd
                col = pd.read_csv(os.path.join(self.logs, 'summary.csv'))['val-PSNR']
                load_epoch = col.idxmax() if load_epoch == 'best' else len(col) - 1
            self.model_epoch = load_epoch
            self.model.load_model(model_save_name='train_model', model_idx=load_epoch,
                                  legacy=self.model.legacy_load, load_override=load_override)
        else:
            self.model.pre_training_model_load()

        self.full_name = '%s_%d' % (experiment, self.model_epoch)
        if gpu == 'multi':
            self.model.set_multi_gpu()
        self.configuration = {'input': self.model.im_input, 'colorspace': self.model.colorspace}
        self.print_overview()

    # -- training / evaluation entry points --------------------------------------------------
    def train_batch(self, lr, hr, **kwargs):
        return self.model.run_train(x=lr, y=hr, **kwargs)

    def set_epoch(self, epoch):
        self.model_epoch = epoch
        self.model.set_epoch(epoch)

    def net_run_and_process(self, lr=None, hr=None, **kwargs):
        """run_eval + clamp + colour conversion (ref :138-169).

        Two optional keywords beyond the reference's, both off by default so that the return values
        stay the reference's fp32 arrays: `output_dtype='uint8'` returns rint(255 * .) arrays (3 B instead of 12 B per
        output pixel and array over PCIe), `device_psnr=True` (needs `hr`) leaves the per-image Y-channel PSNR of
        sr_tools.metrics.psnr(max_value=1) in `self.last_y_psnr` — computed on the GPU from the fp32 result."""
        self.last_y_psnr = None
        want_u8 = kwargs.pop('output_dtype', None) in ('uint8', np.uint8, torch.uint8)
        want_psnr = bool(kwargs.pop('device_psnr', False))
        if 'rgb' in self.configuration['colorspace']:
            fused = self._device_postprocess(lr, hr, kwargs, want_u8=want_u8, want_psnr=want_psnr)
            if fused is not None:
                return fused
            out_rgb, loss, timing = self.model.run_eval(x=lr, y=hr, **kwargs)
            out_ycbcr = self.colorspace_convert(out_rgb, colorspace='rgb')
            out_rgb = self._standard_image_formatting(out_rgb.numpy())
        else:
            f_ref = hr if hr is None else hr[:, 0, :, :].unsqueeze(1)
            out_y, loss, timing = self.model.run_eval(lr[:, 0, :, :].unsqueeze(1), y=f_ref, **kwargs)
            out_ycbcr = torch.stack([out_y.squeeze(1), lr[:, 1, :, :], lr[:, 2, :, :]], 1)
            out_rgb = self.colorspace_convert(out_ycbcr, colorspace='ycbcr')
            out_ycbcr = self._standard_image_formatting(out_ycbcr.numpy())
        return out_rgb, out_ycbcr, loss, timing

    def _device_postprocess(self, lr, hr, kwargs, want_u8=False, want_psnr=False):
        """clip + RGB -> YCbCr on the GPU right after the network: the SR batch stays on the
        device, one kernel writes both result arrays (bit-identical to the numpy expressions below), and they come
        back through pooled page-locked buffers.  Returns None when the model does not keep results on a CUDA device
        (CPU handlers, chopped Q-SAN evaluation) — the caller then takes the reference's host path."""
        if kwargs.get('keep_on_device') or not isinstance(getattr(self.model, 'device', None), int):
            return None
        if type(self.model).run_eval.__qualname__.split('.')[0] not in ('QModel', 'BaseModel'):
            return None  # handlers with their own evaluation protocol (chopped Q-SAN) keep the host path
        out, loss, timing = self.model.run_eval(x=lr, y=hr, keep_on_device=True, **kwargs)
        if not (torch.is_tensor(out) and out.is_cuda and out.dim() == 4 and out.shape[1] == 3
                and out.dtype == torch.float32):
            return self._host_postprocess(out.cpu() if torch.is_tensor(out) and out.is_cuda else out, loss, timing)
        import ctypes as C
        from deepfir_b200 import _lib
        lib = _lib.load_library()
        out = out.contiguous()
        if want_u8 or (want_psnr and hr is not None):
            hw = out.shape[2] * out.shape[3]
            rgb8 = torch.empty(out.shape, dtype=torch.uint8, device=out.device) if want_u8 else None
            ycc8 = torch.empty(out.shape, dtype=torch.uint8, device=out.device) if want_u8 else None
            hr_d = psnr = scratch = None
            if want_psnr and hr is not None:
                hr_d = hr.to(device=out.device, dtype=torch.float32).contiguous()
                if hr_d.shape != out.shape:
                    raise RuntimeError("device_psnr: hr has shape %s, the SR batch %s" % (tuple(hr_d.shape), tuple(out.shape)))
                psnr = torch.empty(out.shape[0], dtype=torch.float32, device=out.device)
                scratch = torch.empty(lib.dfir_postprocess_u8_scratch_bytes(out.shape[0], hw), dtype=torch.uint8,
                                      device=out.device)
            ptr = lambda t: (t.data_ptr() if t is not None else None)
            with torch.cuda.device(out.device):
                _lib.check(lib.dfir_postprocess_u8(out.data_ptr(), ptr(hr_d), ptr(rgb8), ptr(ycc8), ptr(psnr), ptr(scratch),
                                                   scratch.numel() if scratch is not None else 0, out.shape[0], hw,
                                                   C.c_void_p(torch.cuda.current_stream(out.device).cuda_stream)),
                           "postprocess_u8")
            if psnr is not None:
                self.last_y_psnr = psnr.cpu().numpy()
            if want_u8:
                return BaseModel._to_host(rgb8).numpy(), BaseModel._to_host(ycc8).numpy(), loss, timing
        rgb, ycc = torch.empty_like(out), torch.empty_like(out)
        with torch.cuda.device(out.device):
            _lib.check(lib.dfir_postprocess_rgb(out.data_ptr(), rgb.data_ptr(), ycc.data_ptr(), out.shape[0],
                                                out.shape[2] * out.shape[3], 0.0, 1.0,
                                                C.c_void_p(torch.cuda.current_stream(out.device).cuda_stream)),
                       "postprocess_rgb")
        return BaseModel._to_host(rgb).numpy(), BaseModel._to_host(ycc).numpy(), loss, timing

    def _host_postprocess(self, out_rgb, loss, timing):
        out_ycbcr = self.colorspace_convert(out_rgb, colorspace='rgb')
        return self._standard_image_formatting(out_rgb.numpy()), out_ycbcr, loss, timing

    @staticmethod
    def colorspace_convert(image, colorspace='rgb'):
        batch = ModelInterface._standard_image_formatting(image.numpy())
        for i in range(batch.shape[0]):
            batch[i, ...] = ycbcr_convert(batch[i, ...], im_type='jpg', input=colorspace, y_only=False)
        return batch

    @staticmethod
    def _standard_image_formatting(im, min_value=0, max_value=1):
        return np.clip(np.copy(im), min_value, max_value)

    def net_forensic(self, data, **kwargs):
        image, forensic_data = self.model.run_forensic(data, **kwargs)
        return image.numpy(), forensic_data

    # -- persistence -------------------------------------------------------------------------
    def save(self, name='train_model', override=False, dry_run=False):
        save_path = os.path.join(self.saved_models, "{}_{}".format(name, str(self.model_epoch)))
        if os.path.isfile(save_path) and not override:
            raise RuntimeError('Saving this model will result in overwriting existing data!  '
                               'Change model location or enable override.')
        if dry_run:
            print('Training cleared to run.')
        else:
            self.model.save_model(model_save_name=name, model_idx=self.model_epoch)

    def save_metadata(self):
        import pandas as pd
        pd.DataFrame.from_dict({'model_parameters': [self.model.print_parameters()]}).to_csv(
            os.path.join(self.base_folder, 'extra_metadata.csv'), index=False)

    def print_overview(self):
        evaluating = self.mode == 'eval'
        epoch = self.model_epoch if (evaluating or self.model_epoch == 0) else self.model_epoch + 1
        device = self.model.device if str(self.model.device) == 'cpu' else 'GPU ' + str(self.model.device)
        print('----------------------------')
        print('Handler for experiment %s initialized successfully.' % self.experiment)
        print('System loaded in %s mode - %s architecture provided.' % ('eval' if evaluating else 'train', self.name))
        print('Model has %d trainable parameters.' % self.model.print_parameters())
        print("Using %s as the model's primary device, and %s epoch %d of the model."
              % (device, 'currently evaluating' if evaluating else 'will start training from', epoch))
        self.model.extra_diagnostics()
        print('----------------------------')

    @staticmethod
    def define_model(name, **kwargs):
        return locate(available_models[name])(**kwargs)

    @staticmethod
    def _legacy_model_setup(experiment, exp_folder, scale):
        try:
            l_data = read_metadata(os.path.join(exp_folder, 'meta_data.csv'))
        except Exception:
            raise RuntimeError('No metadata information provided - model structure unknown.')
        return {'name': l_data['model'], 'internal_params': {'scale': scale}}

    def epoch_end_calls(self):
        self.model.epoch_end_calls()

    def get_learning_rate(self):
        return self.model.get_learning_rate()


class BaseModel(nn.Module):
    """Common handler functionality (ref :257-574)."""

    def __init__(self, device, model_save_dir, eval_mode, grad_clip=None, **kwargs):
        super(BaseModel, self).__init__()
        self.criterion = nn.L1Loss()
        self.device = torch.device('cpu') if device == 'cpu' else device
        self.optimizer = None
        self.net = None
        self.face_finder = False
        self.model_name = None
        self.im_input = None
        self.colorspace = None
        self.grad_clip = None if grad_clip == 0 else grad_clip
        self.model_save_dir = model_save_dir
        self.eval_mode = eval_mode
        self.curr_epoch = 0
        self.state = {}
        self.learning_rate_scheduler = None
        self.legacy_load = True

    # -- optimisation setup ------------------------------------------------------------------
    def define_optimizer(self, lr=1e-4, optimizer_params=None):
        params = [p for p in self.net.parameters() if p.requires_grad]
        # same Adam as the reference (:291-299); on CUDA parameters the single-kernel ("fused") implementation is
        # selected: the default per-tensor loop costs ~10 ms per step over Q-RCAN's 1648 parameter tensors
        # same Adam as the reference (:291-299).  On CUDA parameters: `FlatAdam`, a torch.optim.Adam subclass whose step is
        # one kernel over flat buffers (the per-tensor implementations cost 2-10 ms per step over Q-RCAN's 1648 tensors)
        if params and all(p.is_cuda for p in params):
            from deepfir_b200.flat_adam import FlatAdam as adam_cls
        else:
            adam_cls = optim.Adam
        if optimizer_params is not None:
            self.optimizer = adam_cls(params, lr=lr, betas=(optimizer_params['beta_1'], optimizer_params['beta_2']))
        else:
            self.optimizer = adam_cls(params, lr=lr)

    def define_scheduler(self, scheduler, scheduler_params):
        sched = optim.lr_scheduler
        if scheduler == 'cosine_annealing_warm_restarts':
            self.learning_rate_scheduler = sched.CosineAnnealingWarmRestarts(
                self.optimizer, T_mult=scheduler_params['t_mult'], T_0=scheduler_params['restart_period'],
                eta_min=scheduler_params['lr_min'])
        elif scheduler == 'multi_step_lr':
            self.learning_rate_scheduler = sched.MultiStepLR(self.optimizer, milestones=scheduler_params['milestones'],
                                                             gamma=scheduler_params['gamma'])
        elif scheduler == 'custom_dasr':
            def dasr(epoch):
                if epoch < 60:
                    return 1e-3
                if epoch < 225:
                    return 1e-4
                return 1e-4 * math.pow(0.5, (epoch - 100) // 125)
            self.learning_rate_scheduler = sched.LambdaLR(self.optimizer, lr_lambda=dasr)
        elif scheduler == 'step_lr':
            self.learning_rate_scheduler = sched.StepLR(self.optimizer, step_size=scheduler_params['step_size'],
                                                        gamma=scheduler_params['gamma'])
        else:
            raise RuntimeError('%s scheduler not implemented' % scheduler)

    def activate_device(self):
        self.net.to(self.device)

    def training_setup(self, lr, scheduler, scheduler_params, perceptual, device, optimizer_params=None):
        if not self.eval_mode:
            self.define_optimizer(lr=lr, optimizer_params=optimizer_params)
            if scheduler is not None:
                self.define_scheduler(scheduler=scheduler, scheduler_params=scheduler_params)
        if perceptual is not None and self.eval_mode is False:
            raise NotImplementedError('perceptual loss (VGG feature extractors) is outside the B200 hot path')

    def set_multi_gpu(self, device_ids=None):
        """The reference wraps the network in nn.DataParallel (models/__init__.py:307).  The B200 networks keep device
        pointer tables, a workspace and a flat gradient buffer per device and cannot be replicated by DataParallel's
        shallow module copies, so more than one device is refused here: multi-GPU runs are one process per GPU
        (torchrun; inference sharded by image with deepfir_b200.sharding, training with the NCCL gradient all-reduce of
        QModel.run_train — see INTEGRATION.md).  A single device keeps the reference's wrapper-free behaviour."""
        ids = list(device_ids) if device_ids is not None else list(range(torch.cuda.device_count()))
        if len(ids) > 1:
            raise NotImplementedError(
                "gpu='multi' (nn.DataParallel) is not supported by the B200 path: launch one process per GPU "
                "(python -m torch.distributed.run --nproc-per-node N ...); see INTEGRATION.md, 'Multi-GPU'")
        if ids and isinstance(self.device, int) and ids[0] != self.device:
            self.device = ids[0]
            self.net.to(self.device)

    # -- checkpoints -------------------------------------------------------------------------
    def save_model(self, model_save_name, model_idx, extract_state_only=False):
        net = self.net.module if isinstance(self.net, nn.DataParallel) else self.net
        self.state['network'] = net.state_dict()
        self.state['optimizer'] = self.optimizer.state_dict()
        self.state['model_name'] = self.model_name
        self.state['model_epoch'] = self.curr_epoch
        if self.learning_rate_scheduler is not None:
            self.state['scheduler_G'] = self.learning_rate_scheduler.state_dict()
        if extract_state_only:
            return self.state
        torch.save(self.state, f=os.path.join(self.model_save_dir, "{}_{}".format(model_save_name, str(model_idx))))

    @staticmethod
    def legacy_switch(state_dict):
        renamed = OrderedDict()
        for k, v in state_dict.items():
            for prefix in ('model.module.', 'model.'):
                if k.startswith(prefix):
                    k = k[len(prefix):]
                    break
            renamed[k] = v
        return renamed

    def load_model(self, model_save_name, model_idx, legacy=False, load_override=None, preloaded_state=None):
        loc = self.device if self.device == torch.device('cpu') else "cuda:%d" % self.device
        folder = self.model_save_dir if load_override is None else load_override
        load_file = os.path.join(folder, "{}_{}".format(model_save_name, str(model_idx)))
        state = torch.load(f=load_file, map_location=loc, weights_only=False) if preloaded_state is None \
            else preloaded_state
        net_state = self.legacy_switch(state['network']) if legacy else state['network']
        self.net.load_state_dict(state_dict=net_state)
        if not self.eval_mode:
            self.optimizer.load_state_dict(state['optimizer'])
            if self.learning_rate_scheduler is not None:
                self.learning_rate_scheduler.load_state_dict(state['scheduler_G'])
        self.set_epoch(state['model_epoch'])
        if state['model_name'] == 'qpircan':
            state['model_name'] = 'qrcan'
        print('Loaded model uses the following architecture:', state['model_name'])
        return state

    # -- execution ---------------------------------------------------------------------------
    def run_train(self, x, y, tag=None, mask=None, keep_on_device=False, *args, **kwargs):
        if self.eval_mode:
            raise RuntimeError('Model initialized in eval mode, training not possible.')
        if not self.net.training:  # (walking ~2600 sub-modules costs milliseconds: only when the mode really changes)
            self.net.train()
        x, y = x.to(device=self.device), y.to(device=self.device)
        out = self.run_model(x, image_names=tag, **kwargs)
        loss = self.criterion(out, y)
        self.standard_update(loss)
        if keep_on_device:
            return loss.detach().cpu().numpy(), out.detach()
        return loss.detach().cpu().numpy(), out.detach().cpu()

    def standard_update(self, loss):
        self.optimizer.zero_grad()
        loss.backward()
        if self.grad_clip is not None:
            nn.utils.clip_grad_norm_(self.net.parameters(), self.grad_clip)
        self.optimizer.step()
        if self.learning_rate_scheduler is not None:
            self.learning_rate_scheduler.step()  # per batch, as in the reference (:488-489)

    # run_eval returns a HOST tensor.  Opt-in (set to a batch size): with at least this many images the batch runs as two
    # halves and the device -> host copy of the first half overlaps the forward of the second (an image's result does not
    # depend on the rest of its batch, so the values are bit-identical).  OFF by default: measured on B200 at 32 x 128x128 the
    # copy that is hidden (0.9 of 1.8 ms) costs less than the second pass adds (421 more launches with their fill / drain:
    # 327 -> 315 MPix/s end to end).  Worth it only where a forward is short against its result copy.
    overlap_d2h_min_batch = None

    def run_eval(self, x, y=None, request_loss=False, tag=None, timing=False, keep_on_device=False, *args, **kwargs):
        if self.net.training:
            self.net.eval()
        elapsed = None
        with torch.no_grad():
            x = x.to(device=self.device)
            halves = None
            if not keep_on_device and not (request_loss and y is not None) and x.is_cuda:
                halves = self._split_batch(x, tag, kwargs)
            if halves is not None:
                if timing:
                    tic = time.perf_counter()
                host = self._run_halves(halves)
                if timing:
                    elapsed = time.perf_counter() - tic
                return host, None, elapsed
            if timing:
                tic = time.perf_counter()
            out = self.run_model(x, image_names=tag, **kwargs)
            if timing:
                elapsed = time.perf_counter() - tic  # launch time only, like the reference (no device sync)
            if request_loss and y is not None:
                loss = self.criterion(out, y.to(device=self.device)).detach().cpu().numpy()
            else:
                loss = None
        out = out.detach()
        return (out if keep_on_device else self._to_host(out)), loss, elapsed

    def _split_batch(self, x, tag, kwargs):
        """[(x_half, tag_half, kwargs_half)] x 2 when the batch is large enough to overlap copy and compute, else None;
        tensors / lists in kwargs whose first dimension is the batch are split with it"""
        n = x.shape[0]
        if self.overlap_d2h_min_batch is None or n < self.overlap_d2h_min_batch or n % 2:
            return None
        h = n // 2
        parts = []
        for sl in (slice(0, h), slice(h, n)):
            kw = {}
            for k, v in kwargs.items():
                if torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == n:
                    kw[k] = v[sl]
                elif isinstance(v, (list, tuple)) and len(v) == n:
                    kw[k] = v[sl]
                else:
                    kw[k] = v
            t = tag[sl] if isinstance(tag, (list, tuple)) and len(tag) == n else tag
            parts.append((x[sl], t, kw))
        return parts

    _COPY_STREAMS = {}

    def _run_halves(self, halves):
        dev = halves[0][0].device
        side = BaseModel._COPY_STREAMS.get(dev)
        if side is None:
            side = BaseModel._COPY_STREAMS[dev] = torch.cuda.Stream(device=dev)
        main = torch.cuda.current_stream(dev)
        xa, ta, kwa = halves[0]
        out_a = self.run_model(xa, image_names=ta, **kwa).detach()
        ev = torch.cuda.Event()
        ev.record(main)
        n = xa.shape[0] + halves[1][0].shape[0]
        host = self._host_buffer((n,) + tuple(out_a.shape[1:]), out_a.dtype)
        side.wait_event(ev)
        with torch.cuda.stream(side):
            host[:xa.shape[0]].copy_(out_a, non_blocking=True)
        xb, tb, kwb = halves[1]
        out_b = self.run_model(xb, image_names=tb, **kwb).detach()
        host[xa.shape[0]:].copy_(out_b, non_blocking=True)
        main.synchronize()
        side.synchronize()
        return host

    _PINNED = {}  # (shape, dtype) -> page-locked result buffers owned by this module

    @classmethod
    def _host_buffer(cls, shape, dtype):
        """a pooled page-locked buffer of this shape.  Page-locking 100 MB costs ~12 ms per call (measured: 7 GB/s through a
        fresh buffer against 56 GB/s into an existing one), so result buffers are pooled and handed out again once the
        caller has dropped every reference to the previous result (storage use count)."""
        key = (tuple(shape), dtype)
        pool = cls._PINNED.setdefault(key, [])
        host = None
        try:
            for buf in pool:  # count 2 = the pool's tensor + the temporary storage wrapper: nobody else looks at it
                if torch._C._storage_Use_Count(buf.untyped_storage()._cdata) <= 2:
                    host = buf
                    break
        except Exception:
            pool = None
        if host is None:
            host = torch.empty(shape, dtype=dtype, pin_memory=True)
            if pool is not None and len(pool) < 4 and len(cls._PINNED) <= 8:
                pool.append(host)
        return host.view(host.shape)  # a new tensor object on the pooled storage; dropping it frees the buffer for re-use

    @classmethod
    def _to_host(cls, t):
        """device -> host like `.cpu()`, but through pooled page-locked memory: the SR batch is ~12 B per output pixel and a
        pageable copy would dominate the end-to-end time."""
        if not t.is_cuda:
            return t
        out = cls._host_buffer(tuple(t.shape), t.dtype)
        out.copy_(t, non_blocking=True)
        torch.cuda.current_stream(t.device).synchronize()
        return out

    def run_forensic(self, x, *args, **kwargs):
        self.net.eval()
        with torch.no_grad():
            out, data = self.net.forensic(x.to(device=self.device), **kwargs)
        return out.cpu().detach(), data

    def run_model(self, x, *args, **kwargs):
        return self.net.forward(x)

    def print_parameters(self, verbose=False):
        total = 0
        for name, value in self.named_parameters():
            if verbose:
                print(name, value.shape)
            total += np.prod(value.shape)
        if verbose:
            print('Total number of trainable parameters:', total)
        return total

    def print_status(self):
        raise NotImplementedError

    def epoch_end_calls(self):
        pass

    def set_epoch(self, epoch):
        self.curr_epoch = epoch

    def get_learning_rate(self):
        return self.optimizer.param_groups[0]['lr']

    def extra_diagnostics(self):
        pass

    def pre_training_model_load(self):
        pass
