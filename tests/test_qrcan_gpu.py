"""-m gpu: whole-network parity of the B200 Q-RCAN path against the reference's golden outputs / the oracle."""
import pytest
import torch

from oracle import deepfir_oracle as O
from tests.golden_util import case_tensors, golden_names, load_golden, max_norm_err, oracle_forward

pytestmark = pytest.mark.gpu

QRCAN_CASES = [n for n in golden_names() if n.startswith("qrcan") and "pa_" not in n]


def _build(info, precision):
    from deepfir_b200.qrcan import QRCAN
    net = QRCAN(precision=precision, **info["kwargs"])
    sd, x, meta = case_tensors(info)
    net.load_state_dict(sd, strict=True)
    return net.cuda().eval(), x, meta


@pytest.mark.parametrize("name", QRCAN_CASES)
def test_qrcan_fp32_mode_matches_reference_golden(name):
    """north_star fp32 mode: max|out-ref| / max|ref| <= 1e-4 (SURVEY.md §8d tolerance)."""
    ref, info = load_golden(name)
    net, x, meta = _build(info, "fp32")
    with torch.no_grad():
        out = net(x.cuda(), meta.cuda()).cpu()
    assert out.shape == ref.shape
    assert max_norm_err(out, ref) <= 1e-4


@pytest.mark.parametrize("name", QRCAN_CASES)
def test_qrcan_bf16_mode_matches_reference_golden(name):
    """north_star bf16 mode: SR-output PSNR delta within 0.01 dB.  With random-init weights there is no
    meaningful HR target, so the check is two-fold: (1) against a synthetic HR target built from the
    reference output + noise at ~30 dB, |PSNR(out,HR) - PSNR(ref,HR)| <= 0.01 dB; (2) PSNR(out, ref)
    >= 56.4 dB, the bound that implies (1) for any ~30 dB reconstruction (SURVEY.md §7 hard part 5)."""
    ref, info = load_golden(name)
    net, x, meta = _build(info, "bf16")
    with torch.no_grad():
        out = net(x.cuda(), meta.cuda()).cpu()
    assert out.shape == ref.shape and torch.isfinite(out).all()
    g = torch.Generator().manual_seed(0)
    hr = (ref + torch.randn(ref.shape, generator=g) * 10 ** (-30 / 20)).clamp(0, 1)
    d = abs(O.psnr(out.clamp(0, 1), hr) - O.psnr(ref.clamp(0, 1), hr))
    assert d <= 0.01, d
    assert O.psnr(out, ref, max_value=1.0) >= 56.4


def test_qrcan_bf16_full_depth_matches_bf16_policy_oracle():
    """10x20 RCAB network: the GPU path must agree with the oracle run under the same rounding policy
    (bf16 conv operands, fp32 everything else) far more tightly than with the exact fp32 oracle."""
    ref, info = load_golden("qrcan_standard_full")
    net, x, meta = _build(info, "bf16")
    sd, _, _ = case_tensors(info)
    with torch.no_grad():
        out = net(x.cuda(), meta.cuda()).cpu()
        pol = oracle_forward(info, sd, x, meta, nm=O.Numerics(torch.bfloat16))
    assert max_norm_err(out, pol) < 5e-3
    assert max_norm_err(out, ref) < 3e-2


def test_batch_composition_does_not_change_an_image():
    """sharded inference == unsharded, bit-exact per image: per-row pooled sums make every image's result
    independent of what else is in the batch."""
    ref, info = load_golden("qrcan_standard_g2b2")
    net, x, meta = _build(info, "bf16")
    xs = torch.cat([x, x.flip(0), x], 0).cuda()
    ms = torch.cat([meta, meta.flip(0), meta], 0).cuda()
    with torch.no_grad():
        full = net(xs, ms)
        one = net(xs[1:2].contiguous(), ms[1:2].contiguous())
    assert torch.equal(full[1:2], one)
    assert torch.equal(full[0], full[4])


def test_state_dict_roundtrip_and_repack():
    """weights are re-packed when parameters change (the packed tiles are a cache, never saved)."""
    ref, info = load_golden("qrcan_noq_scale2")
    net, x, meta = _build(info, "bf16")
    with torch.no_grad():
        a = net(x.cuda(), meta.cuda())
        for p in net.parameters():
            p.mul_(0.5)
        b = net(x.cuda(), meta.cuda())
        sd, _, _ = case_tensors(info)
        net.load_state_dict(sd)
        c = net(x.cuda(), meta.cuda())
    assert not torch.equal(a, b)
    assert torch.equal(a, c)
