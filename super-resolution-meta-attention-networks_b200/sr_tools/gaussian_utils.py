"""Online degradation for training (mirror of the reference's Code/sr_tools/gaussian_utils.py:196-424, the part the
training input pipeline uses): Gaussian blur-kernel generators (host, numpy — a few hundred numbers per image), and
BatchBlur / PCAEncoder / SRMDPreprocessing whose tensor work runs on the GPU through libdfir_b200.so
(dfir_batch_blur, dfir_pca_encode).  There is no CPU path: `cuda=False` raises (the reference's CPU implementation is the
thing being replaced; tests compare against a restatement of it in oracle/)."""
import ctypes as C
import math

import numpy as np
import torch

from deepfir_b200 import _lib


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


# ------------------------------------------------------------------ blur kernels (reference :203-273)
def _grid(l):
    ax = np.arange(-l // 2 + 1., l // 2 + 1.)
    return np.meshgrid(ax, ax)


def isotropic_gaussian_kernel(l, sigma):
    xx, yy = _grid(l)
    k = np.exp(-(xx ** 2 + yy ** 2) / (2. * sigma ** 2))
    return k / k.sum()


def anisotropic_gaussian_kernel(l, sig_x, sig_y, radians):
    """Gaussian with covariance R diag(sig_x^2, sig_y^2) R^T (reference cal_sigma + anisotropic_gaussian_kernel)"""
    c, s = math.cos(radians), math.sin(radians)
    rot = np.array([[c, -s], [s, c]])
    inv = np.linalg.inv(rot @ np.diag([sig_x ** 2, sig_y ** 2]) @ rot.T)
    xx, yy = _grid(l)
    xy = np.stack([xx, yy], axis=-1)
    k = np.exp(-0.5 * np.einsum("hwi,ij,hwj->hw", xy, inv, xy))
    return k / k.sum()


def random_gaussian_kernel(l=21, sig_min=0.2, sig_max=4.0, rate_iso=1.0, scaling=3, rng=np.random):
    """same sequence of draws as the reference (:226-250): iso/aniso choice, then the widths (and the angle first for
    anisotropic kernels)"""
    if rng.random() < rate_iso:
        return isotropic_gaussian_kernel(l, rng.random() * (sig_max - sig_min) + sig_min)
    angle = rng.random() * math.pi * 2 - math.pi
    sx = rng.random() * (sig_max - sig_min) + sig_min
    sy = float(np.clip(rng.random() * scaling * sx, sig_min, sig_max))
    return anisotropic_gaussian_kernel(l, sx, sy, angle)


class BatchSRKernel(object):
    """reference :318-331: a batch of random kernels, or `batch` copies of the stable isotropic one"""

    def __init__(self, l=21, sig=2.6, sig_min=0.2, sig_max=4.0, rate_iso=1.0, scaling=3):
        self.l, self.sig, self.sig_min, self.sig_max, self.rate, self.scaling = l, sig, sig_min, sig_max, rate_iso, scaling

    def __call__(self, random, batch, tensor=False):
        if random:
            ks = np.stack([random_gaussian_kernel(self.l, self.sig_min, self.sig_max, self.rate, self.scaling)
                           for _ in range(batch)])
        else:
            ks = np.stack([isotropic_gaussian_kernel(self.l, self.sig)] * batch)
        return torch.FloatTensor(ks) if tensor else ks


def random_batch_noise(batch, high, rate_cln=1.0):
    level = np.random.uniform(size=(batch, 1)) * high
    mask = (np.random.uniform(size=(batch, 1)) >= rate_cln).astype(level.dtype)
    return level * mask


# ------------------------------------------------------------------ tensor work on the GPU
def _need_cuda(t, what):
    if not (torch.is_tensor(t) and t.is_cuda):
        raise RuntimeError("%s runs on a CUDA (sm_100a) device only: there is no CPU path" % what)


class PCAEncoder(object):
    """reference :333-343: kernel code = flattened kernel @ PCA matrix ([l*l, k])"""

    def __init__(self, weight, cuda=False):
        if not cuda:
            raise RuntimeError("PCAEncoder: only cuda=True is available on the B200 path")
        self.weight = weight.detach().to(device="cuda", dtype=torch.float32).contiguous()
        self.size = self.weight.size()

    def __call__(self, batch_kernel, noise_sigma=None):
        _need_cuda(batch_kernel, "PCAEncoder")
        B, H, W = batch_kernel.size()
        k = self.size[1]
        bk = batch_kernel.to(torch.float32).contiguous()
        sg = noise_sigma.reshape(-1).to(device=bk.device, dtype=torch.float32).contiguous() if noise_sigma is not None else None
        code = torch.empty(B, k + (1 if sg is not None else 0), device=bk.device, dtype=torch.float32)
        with torch.cuda.device(bk.device):
            _lib.check(_lib.load_library().dfir_pca_encode(bk.data_ptr(), self.weight.to(bk.device).data_ptr(),
                                                           sg.data_ptr() if sg is not None else None, code.data_ptr(), B, H, k,
                                                           _stream(bk.device)), "pca_encode")
        return code


class BatchBlur(torch.nn.Module):
    """reference :346-368: reflection pad + one l x l kernel per image (3-D `kernel`) or one shared kernel (2-D)"""

    def __init__(self, l=15):
        super().__init__()
        self.l = l

    def forward(self, input, kernel, noise=None, noise_sigma=None, clamp01=False):
        _need_cuda(input, "BatchBlur")
        B, Cc, H, W = input.size()
        x = input.to(torch.float32).contiguous()
        kern = kernel.to(device=x.device, dtype=torch.float32).contiguous()
        out = torch.empty_like(x)
        ptr = lambda t: (t.data_ptr() if t is not None else None)
        if noise is not None:
            noise = noise.to(device=x.device, dtype=torch.float32).contiguous()
            noise_sigma = noise_sigma.reshape(-1).to(device=x.device, dtype=torch.float32).contiguous()
        with torch.cuda.device(x.device):
            _lib.check(_lib.load_library().dfir_batch_blur(x.data_ptr(), kern.data_ptr(), 1 if kern.dim() == 3 else 0,
                                                           ptr(noise), ptr(noise_sigma), out.data_ptr(), B, Cc, H, W, self.l,
                                                           1 if clamp01 else 0, _stream(x.device)), "batch_blur")
        return out


class SRMDPreprocessing(object):
    """reference :371-424: random (or stable) blur kernel per image -> blur -> Gaussian noise + clamp -> kernel code
    (+ 10 * noise level).  Blur, noise application, clamp and PCA projection are two kernel launches on the GPU; the kernel
    parameters and noise levels are drawn on the host exactly as in the reference, the noise samples on the device."""

    def __init__(self, pca, random, para_input=10, kernel=21, noise=True, cuda=False, sig=2.6, sig_min=0.2, sig_max=4.0,
                 rate_iso=1.0, scaling=3, rate_cln=0.2, noise_high=0.08, **kwargs):
        if not cuda:
            raise RuntimeError("SRMDPreprocessing: only cuda=True is available on the B200 path")
        self.encoder = PCAEncoder(pca, cuda=True)
        self.kernel_gen = BatchSRKernel(l=kernel, sig=2.6 if sig is None else sig, sig_min=sig_min, sig_max=sig_max,
                                        rate_iso=rate_iso, scaling=scaling)
        self.blur = BatchBlur(l=kernel)
        self.para_in, self.l, self.noise, self.cuda = para_input, kernel, noise, True
        self.rate_cln, self.noise_high, self.random = rate_cln, noise_high, random

    def __call__(self, hr_tensor, generator=None):
        """hr_tensor: [C,H,W] (like the reference) or [B,C,H,W]; returns (lr_re, re_code, b_kernels) on the device"""
        if hr_tensor.dim() == 3:
            hr_tensor = hr_tensor.unsqueeze(0)
        hr = hr_tensor.to(device="cuda", dtype=torch.float32)
        B = hr.shape[0]
        b_kernels = self.kernel_gen(self.random, B, tensor=True).to(hr.device)
        if self.noise:
            level = torch.from_numpy(random_batch_noise(B, self.noise_high, self.rate_cln)).float().to(hr.device)
            samples = torch.randn(hr.shape, device=hr.device, generator=generator)
            lr = self.blur(hr, b_kernels, noise=samples, noise_sigma=level, clamp01=True)
            code = self.encoder(b_kernels, noise_sigma=level)
        else:
            lr = self.blur(hr, b_kernels)
            code = self.encoder(b_kernels)
        return lr, code, b_kernels
