"""Pins the CPU oracle (oracle/deepfir_oracle.py) against outputs of the reference itself.

The reference has no tests or golden vectors for this path; the fixtures in
tests/golden/ were produced by oracle/make_golden.py from the live reference."""
import numpy as np
import pytest
import torch

from oracle import deepfir_oracle as O
from oracle import np_ops
from oracle.ref_shim import reference_available
from tests.golden_util import case_tensors, golden_names, load_golden, max_norm_err, oracle_forward


@pytest.mark.parametrize("name", golden_names())
def test_oracle_matches_reference_golden(name):
    if name == "qrcan_standard_full":
        torch.set_num_threads(max(1, torch.get_num_threads()))
    out_ref, info = load_golden(name)
    sd, x, meta = case_tensors(info)
    with torch.no_grad():
        out = oracle_forward(info, sd, x, meta)
    assert out.shape == out_ref.shape
    # fp32 vs fp32 with a different summation order only (the reference's own fp32-vs-fp64 noise
    # floor on this metric is 7.9e-7)
    assert max_norm_err(out, out_ref) < 2e-5


def test_primitives_against_numpy_float64():
    rs = np.random.RandomState(0)
    x = rs.randn(2, 5, 7, 6)
    w = rs.randn(8, 5, 3, 3)
    b = rs.randn(8)
    sd = {"c.weight": torch.tensor(w), "c.bias": torch.tensor(b)}
    got = O.conv3x3(torch.tensor(x), sd, "c").numpy()
    np.testing.assert_allclose(got, np_ops.conv3x3_np(x, w, b), rtol=1e-10, atol=1e-10)
    for r in (2, 3):
        y = rs.randn(2, 2 * r * r, 4, 5)
        np.testing.assert_array_equal(O.pixel_shuffle(torch.tensor(y), r).numpy(), np_ops.pixel_shuffle_np(y, r))
        np.testing.assert_array_equal(torch.nn.functional.pixel_shuffle(torch.tensor(y), r).numpy(),
                                      np_ops.pixel_shuffle_np(y, r))
    z = rs.randn(2, 6, 5, 4)
    np.testing.assert_allclose(O.covpool(torch.tensor(z)).numpy(), np_ops.covpool_np(z), rtol=1e-9, atol=1e-12)


def test_sqrtm_ns_approximates_matrix_sqrt():
    rs = np.random.RandomState(1)
    a = rs.randn(2, 16, 64)
    cov = torch.tensor(np.einsum("bcm,bdm->bcd", a, a) / 64 + 0.5 * np.eye(16))
    s = O.sqrtm_ns(cov, 5)
    rel = ((s.bmm(s) - cov).norm() / cov.norm()).item()
    assert rel < 0.05  # 5 Newton-Schulz iterations: approximate square root


@pytest.mark.skipif(not reference_available(), reason="live reference only exists in the build container")
@pytest.mark.parametrize("name", ["qrcan_standard_g2b2", "qedsr_f64_b3", "qsan_g2b2", "qhan_b1"])
def test_golden_is_reproducible_from_live_reference(name):
    from oracle.make_golden import run_case
    out_ref, info = load_golden(name)
    shapes, out = run_case(name)
    assert shapes == info["shapes"]
    assert max_norm_err(out, out_ref) < 1e-6


def test_generate_channels_and_scale_qpi_semantics():
    md = torch.tensor([[0.5, 1.0, 2.0], [0.25, 3.0, 4.0]], dtype=torch.float64)
    keys = [("qpi", "qpi"), ("blur_kernel", "blur_kernel"), ("blur_kernel", "blur_kernel")]
    out = O.generate_channels(md, keys, ["blur_kernel"], 2)
    assert out.shape == (2, 2, 1, 1) and out.dtype == torch.float32
    assert out.flatten().tolist() == [1.0, 2.0, 3.0, 4.0]
    g = O.scale_qpi(torch.tensor([[0.5]]).reshape(1, 1, 1, 1))
    assert g.shape == (1, 64, 1, 1)
    mu = 0.5 * 1.0 - 0.2
    base = np.linspace(0, 1, 64)
    want = (1 / (np.sqrt(2 * np.pi) * 0.2)) * np.exp(-(base - mu) ** 2 / (2 * 0.04))
    np.testing.assert_allclose(g.flatten().numpy(), want.astype(np.float32), rtol=1e-6)


@pytest.mark.parametrize("name", ["big_qrcan_full_2x128", "big_qsan_g2b2_128"])
def test_oracle_matches_reference_fingerprint_at_baseline_shape(name):
    """the oracle at the BASELINE.json shapes against fingerprints of the live reference (oracle/make_golden_big.py):
    sampled pixels and projections of the full output"""
    from tests.golden_util import big_fingerprint_errors, load_big_golden
    fp, info = load_big_golden(name)
    sd, x, meta = case_tensors(info)
    with torch.no_grad():
        out = oracle_forward(info, sd, x, meta)
    err_sub, err_proj, _ = big_fingerprint_errors(out, fp, info)
    assert err_sub <= 2e-5 and err_proj <= 2e-5, (err_sub, err_proj)


@pytest.mark.skipif(not reference_available(), reason="reference not mounted")
def test_blur_and_pca_restatements_match_the_live_reference():
    """oracle/np_ops.batch_blur_np / pca_encode_np against the reference's own BatchBlur / PCAEncoder
    (Code/sr_tools/gaussian_utils.py:333-368), odd and even kernel sizes, per-image and shared kernels"""
    import importlib.util
    import os
    from oracle.ref_shim import REFERENCE_CODE
    spec = importlib.util.spec_from_file_location("_ref_gaussian_utils", os.path.join(REFERENCE_CODE, "sr_tools", "gaussian_utils.py"))
    try:
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    except Exception as e:  # optional third-party imports of that file (scipy, PIL, torchvision)
        pytest.skip("reference gaussian_utils not importable here: %r" % (e,))
    g = torch.Generator().manual_seed(2)
    for l, shape in ((21, (2, 3, 30, 41)), (20, (2, 1, 25, 25)), (7, (1, 3, 9, 16))):
        x = torch.rand(*shape, generator=g)
        k = torch.rand(shape[0], l, l, generator=g)
        k = k / k.sum(dim=(1, 2), keepdim=True)
        ref = mod.BatchBlur(l=l)(x, k).numpy()
        assert np.abs(np_ops.batch_blur_np(x.numpy(), k.numpy()) - ref).max() <= 1e-6
        ref1 = mod.BatchBlur(l=l)(x, k[0]).numpy()
        assert np.abs(np_ops.batch_blur_np(x.numpy(), k[0].numpy()) - ref1).max() <= 1e-6
    pca = torch.randn(441, 10, generator=g)
    k = torch.rand(3, 21, 21, generator=g)
    assert np.abs(np_ops.pca_encode_np(k.numpy(), pca.numpy()) - mod.PCAEncoder(pca)(k).numpy()).max() <= 1e-4
