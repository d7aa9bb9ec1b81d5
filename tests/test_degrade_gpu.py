"""-m gpu: the training-input degradation on the GPU against the numpy restatement of the reference's
BatchBlur / PCAEncoder / SRMDPreprocessing (Code/sr_tools/gaussian_utils.py:333-424)."""
import numpy as np
import pytest
import torch

from oracle import np_ops

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape,l,per_image", [((2, 3, 40, 52), 21, True), ((1, 3, 64, 64), 21, False), ((3, 1, 33, 70), 15, True),
                                                ((2, 3, 25, 31), 20, True), ((1, 2, 12, 12), 7, True)])
def test_batch_blur_matches_the_reference_expression(shape, l, per_image):
    from sr_tools.gaussian_utils import BatchBlur, BatchSRKernel
    B, C, H, W = shape
    g = torch.Generator().manual_seed(H + W + l)
    x = torch.rand(B, C, H, W, generator=g)
    np.random.seed(l)
    ks = BatchSRKernel(l=l, rate_iso=0.5)(True, B, tensor=True)
    kern = ks if per_image else ks[0]
    want = np_ops.batch_blur_np(x.numpy(), kern.numpy())
    out = BatchBlur(l)(x.cuda(), kern.cuda()).cpu().numpy()
    assert out.shape == want.shape
    assert np.abs(out - want).max() <= 2e-6
    # a blur kernel sums to one: a constant image stays constant (reflection padding keeps the borders exact)
    const = torch.full((B, C, H, W), 0.375)
    assert np.abs(BatchBlur(l)(const.cuda(), kern.cuda()).cpu().numpy() - 0.375).max() <= 2e-6


def test_pca_encode_and_srmd_preprocessing():
    from sr_tools.gaussian_utils import PCAEncoder, SRMDPreprocessing
    g = torch.Generator().manual_seed(3)
    pca = torch.randn(441, 10, generator=g) * 0.05
    kern = torch.rand(5, 21, 21, generator=g)
    kern = kern / kern.sum(dim=(1, 2), keepdim=True)
    code = PCAEncoder(pca, cuda=True)(kern.cuda()).cpu().numpy()
    assert np.abs(code - np_ops.pca_encode_np(kern.numpy(), pca.numpy())).max() <= 1e-6
    # the whole degradation: blur + noise + clamp + code, against the same steps on the host with the same draws
    np.random.seed(11)
    prep = SRMDPreprocessing(pca, random=True, para_input=10, kernel=21, noise=True, cuda=True, rate_iso=0.3, rate_cln=0.4)
    hr = torch.rand(4, 3, 48, 40, generator=g)
    gen = torch.Generator(device="cuda").manual_seed(5)
    lr, re_code, b_kernels = prep(hr, generator=gen)
    assert lr.shape == hr.shape and re_code.shape == (4, 11) and b_kernels.shape == (4, 21, 21)
    gen2 = torch.Generator(device="cuda").manual_seed(5)
    samples = torch.randn(hr.shape, device="cuda", generator=gen2).cpu().numpy()
    level = re_code[:, 10].cpu().numpy() / 10.0
    want = np.clip(np_ops.batch_blur_np(hr.numpy(), b_kernels.cpu().numpy()) + samples * level.reshape(4, 1, 1, 1), 0.0, 1.0)
    assert np.abs(lr.cpu().numpy() - want).max() <= 3e-6
    assert np.abs(re_code[:, :10].cpu().numpy() - np_ops.pca_encode_np(b_kernels.cpu().numpy(), pca.numpy())).max() <= 1e-6
    assert float(lr.min()) >= 0.0 and float(lr.max()) <= 1.0
    with pytest.raises(RuntimeError):
        SRMDPreprocessing(pca, random=True, cuda=False)
