"""-m gpu: whole-network parity of the B200 Q-RCAN path against the reference's golden outputs / the oracle."""
import pytest
import torch

from oracle import deepfir_oracle as O
from tests.golden_util import case_tensors, golden_names, load_golden, max_norm_err, oracle_forward

pytestmark = pytest.mark.gpu

QRCAN_CASES = [n for n in golden_names() if n.startswith("qrcan")]  # incl. pixel attention + selective q layers


def _build(info, precision, **extra):
    from deepfir_b200.han_san import QHAN, QSAN
    from deepfir_b200.baselines import EDSR, HAN, RCAN, SAN
    from deepfir_b200.qrcan import QEDSR, QRCAN
    cls = {"qedsr": QEDSR, "qrcan": QRCAN, "qsan": QSAN, "qhan": QHAN, "rcan": RCAN, "edsr": EDSR, "san": SAN,
           "han": HAN}[info["model"]]
    net = cls(precision=precision, **extra, **info["kwargs"])
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


@pytest.mark.parametrize("name", ["qedsr_f64_b3", "qedsr_f256_b2_nl"])
def test_qedsr_fp32_mode_matches_reference_golden(name):
    """Q-EDSR (ParamResBlock chain, 64 and 256 features, with/without the meta MLP's ReLU), fp32 mode <= 1e-4."""
    ref, info = load_golden(name)
    net, x, meta = _build(info, "fp32")
    with torch.no_grad():
        out = net(x.cuda(), meta.cuda()).cpu()
    assert max_norm_err(out, ref) <= 1e-4


@pytest.mark.parametrize("schedule", ["linear", "fused", "streamer"])
def test_qedsr_bf16_mode_matches_reference_golden(schedule):
    ref, info = load_golden("qedsr_f64_b3")
    net, x, meta = _build(info, "bf16", schedule=schedule)
    with torch.no_grad():
        out = net(x.cuda(), meta.cuda()).cpu()
    _, pol_err = _policy_error(info, ref)
    assert max_norm_err(out, ref) <= 2.0 * pol_err + 1e-4, (max_norm_err(out, ref), pol_err)


@pytest.mark.parametrize("name", ["qsan_g2b2", "qhan_b1"])
def test_qsan_qhan_fp32_mode_matches_reference_golden(name):
    """Q-SAN (non-local attention, SOCA with covariance pooling + Newton-Schulz) and Q-HAN (LAM, CSAM) against
    the reference's outputs; the zero-initialised branches (gamma, W) are randomised in the fixtures."""
    ref, info = load_golden(name)
    net, x, meta = _build(info, "fp32")
    with torch.no_grad():
        out = net(x.cuda(), meta.cuda()).cpu()
    assert out.shape == ref.shape
    assert max_norm_err(out, ref) <= 1e-4, max_norm_err(out, ref)


@pytest.mark.parametrize("name", ["qsan_g2b2", "qhan_b1"])
def test_qsan_qhan_bf16_mode_matches_reference_golden(name):
    ref, info = load_golden(name)
    net, x, meta = _build(info, "bf16")
    with torch.no_grad():
        out = net(x.cuda(), meta.cuda()).cpu()
    _, pol_err = _policy_error(info, ref)
    assert max_norm_err(out, ref) <= 2.0 * pol_err + 1e-4, (max_norm_err(out, ref), pol_err)


def test_qsan_handler_forward_chop_matches_oracle_chop():
    """QSANHandler.run_eval always chops into 4 overlapping quadrants (ref handlers.py:99-150)."""
    import tempfile
    from SISR.models import ModelInterface
    from oracle.synth import synth_state_dict, synth_inputs
    h = ModelInterface.define_model("qsan", device=0, model_save_dir=tempfile.gettempdir(), eval_mode=True,
                                    metadata=["blur_kernel"], precision="fp32", max_combined_im_size=20000)
    sd = synth_state_dict({k: list(v.shape) for k, v in h.net.state_dict().items()}, seed=3)
    h.net.load_state_dict(sd)
    x, meta = synth_inputs(1, 26, 30, 10, seed=3)
    keys = [("blur_kernel",)] * 10
    out, _, _ = h.run_eval(x, metadata=meta.reshape(1, 10).double(), metadata_keys=keys)
    # oracle: the same chop/stitch around the oracle network
    info = dict(model="qsan", kwargs={})

    def net_fn(q):
        with torch.no_grad():
            return O.qsan_forward(q, meta, sd)
    b, c, hh, ww = x.shape
    hs, ws_ = hh // 2 + 10, ww // 2 + 10
    quads = [x[:, :, :hs, :ws_], x[:, :, :hs, ww - ws_:], x[:, :, hh - hs:, :ws_], x[:, :, hh - hs:, ww - ws_:]]
    sr = [net_fn(q) for q in quads]
    s = 4
    H2, W2, hh2, wh2, hs2, ws2 = s * hh, s * ww, s * (hh // 2), s * (ww // 2), s * hs, s * ws_
    want = torch.empty(b, c, H2, W2)
    want[:, :, :hh2, :wh2] = sr[0][:, :, :hh2, :wh2]
    want[:, :, :hh2, wh2:] = sr[1][:, :, :hh2, ws2 - W2 + wh2:]
    want[:, :, hh2:, :wh2] = sr[2][:, :, hs2 - H2 + hh2:, :wh2]
    want[:, :, hh2:, wh2:] = sr[3][:, :, hs2 - H2 + hh2:, ws2 - W2 + wh2:]
    assert max_norm_err(out.cpu(), want) <= 1e-4


def _policy_error(info, ref):
    """error of the bf16 numerics policy itself, emulated on the CPU oracle: conv operands (activations and
    weights) rounded to bf16, fp32 accumulation, fp32 residual stream / pooling / attention vectors."""
    sd, x, meta = case_tensors(info)
    with torch.no_grad():
        pol = oracle_forward(info, sd, x, meta, nm=O.Numerics(torch.bfloat16))
    return pol, max_norm_err(pol, ref)


@pytest.mark.parametrize("name", QRCAN_CASES)
def test_qrcan_bf16_mode_matches_reference_golden(name):
    """bf16 tensor-core mode on every configuration: the deviation from the reference must be explained by
    the storage-rounding policy alone — at most twice the error the same policy produces when emulated
    in the CPU oracle (rounding decisions differ between two fp32 summation orders, so the two noise
    realisations are independent; a wiring bug shows up as orders of magnitude more)."""
    ref, info = load_golden(name)
    net, x, meta = _build(info, "bf16")
    with torch.no_grad():
        out = net(x.cuda(), meta.cuda()).cpu()
    assert out.shape == ref.shape and torch.isfinite(out).all()
    _, pol_err = _policy_error(info, ref)
    assert max_norm_err(out, ref) <= 2.0 * pol_err + 1e-4, (max_norm_err(out, ref), pol_err)
    assert O.psnr(out, ref, max_value=1.0) >= 50.0


def test_qrcan_bf16_full_depth_psnr_delta():
    """north_star bf16 tolerance on the published 10x20 architecture: SR-output PSNR delta within 0.01 dB.
    No trained weights exist offline, so the HR target is synthetic: the reference output plus noise at
    30 dB (a typical x4 reconstruction quality), both clipped to [0,1] like net_run_and_process does.
    Also reported: PSNR(out, ref) — 56.4 dB or more implies the bound for any <= 30 dB reconstruction."""
    ref, info = load_golden("qrcan_standard_full")
    net, x, meta = _build(info, "bf16")
    with torch.no_grad():
        out = net(x.cuda(), meta.cuda()).cpu()
    g = torch.Generator().manual_seed(0)
    hr = (ref + torch.randn(ref.shape, generator=g) * 10 ** (-30 / 20)).clamp(0, 1)
    d = abs(O.psnr(out.clamp(0, 1), hr) - O.psnr(ref.clamp(0, 1), hr))
    p = O.psnr(out, ref, max_value=1.0)
    print("full depth bf16: PSNR(out, ref) = %.2f dB, PSNR delta vs synthetic 30 dB target = %.4f dB" % (p, d))
    assert d <= 0.01, d
    assert p >= 56.4, p
    pol, pol_err = _policy_error(info, ref)
    assert max_norm_err(out, ref) <= 2.0 * pol_err + 1e-4


@pytest.mark.parametrize("name", ["qrcan_standard_g2b2", "qrcan_extended_scale8", "qrcan_mini_concat",
                                  "qrcan_max_concat_scale3"])
@pytest.mark.parametrize("schedule", ["fused", "streamer"])
def test_alternative_schedules_compute_the_same_network(name, schedule):
    """the three block-chain schedules (pool-by-linearity [default], fused-in, streamer) agree within the bf16
    policy error."""
    ref, info = load_golden(name)
    net_a, x, meta = _build(info, "bf16")
    net_b, _, _ = _build(info, "bf16", schedule=schedule)
    with torch.no_grad():
        a = net_a(x.cuda(), meta.cuda()).cpu()
        b = net_b(x.cuda(), meta.cuda()).cpu()
    _, pol_err = _policy_error(info, ref)
    assert max_norm_err(b, ref) <= 2.0 * pol_err + 1e-4
    assert max_norm_err(a, b) <= 2.0 * pol_err + 1e-4


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


def test_host_results_stay_valid_while_referenced(tmp_path):
    """run_eval returns host tensors in pooled page-locked buffers: a buffer is only handed out again after the caller
    dropped every reference to the previous result, so results held across calls never change"""
    from SISR.models import ModelInterface
    torch.manual_seed(8)
    h = ModelInterface.define_model("qrcan", device=0, model_save_dir=str(tmp_path), eval_mode=True,
                                    metadata=["blur_kernel"], n_resgroups=1, n_resblocks=1, n_feats=64, scale=2,
                                    style="standard", include_q_layer=True)
    g = torch.Generator().manual_seed(1)
    meta = torch.rand(2, 10, generator=g, dtype=torch.float64) * 0.4
    keys = [("blur_kernel",) * 2] * 10
    xs = [torch.rand(2, 3, 16, 16, generator=g) for _ in range(6)]
    held = [h.run_eval(x, metadata=meta, metadata_keys=keys)[0] for x in xs]      # six live results
    copies = [t.clone() for t in held]
    for x in xs:                                                                   # more calls while they are alive
        h.run_eval(x, metadata=meta, metadata_keys=keys)
    assert len({t.data_ptr() for t in held}) == 6
    for t, c in zip(held, copies):
        assert torch.equal(t, c)
    assert all(t.is_pinned() for t in held[:4])


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


def test_qedsr_256_features_on_the_tensor_cores_matches_reference_golden():
    """the published Q-EDSR width (256 features, BASELINE.json configs[2]) as 4 x 4 blocks of the 64-channel tcgen05 conv
    with fp32 accumulation over input chunks (deepfir_b200/wide.py): deviation explained by the bf16 storage policy"""
    ref, info = load_golden("qedsr_f256_b2_nl")
    net, x, meta = _build(info, "bf16")
    with torch.no_grad():
        out = net(x.cuda(), meta.cuda()).cpu()
    assert out.shape == ref.shape and torch.isfinite(out).all()
    _, pol_err = _policy_error(info, ref)
    assert max_norm_err(out, ref) <= 2.0 * pol_err + 1e-4, (max_norm_err(out, ref), pol_err)
    assert O.psnr(out, ref, max_value=1.0) >= 50.0


@pytest.mark.parametrize("feats,scale,shape", [(128, 2, (2, 9, 150)), (192, 3, (1, 7, 20)), (256, 4, (1, 16, 136))])
def test_qedsr_wide_widths_scales_and_ragged_rows_against_oracle(feats, scale, shape):
    """other plane counts (2, 3, 4), upsampler factors and row widths that are not multiples of the 128-pixel tile"""
    from deepfir_b200.qrcan import QEDSR
    torch.manual_seed(3)
    kw = dict(num_blocks=2, num_features=feats, input_para=10, scale=scale, res_scale=0.1, q_layer_nonlinearity=True)
    net = QEDSR(precision="bf16", **kw)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    b, h, w = shape
    g = torch.Generator().manual_seed(4)
    x = torch.rand(b, 3, h, w, generator=g)
    meta = torch.rand(b, 10, 1, 1, generator=g) * 0.4
    with torch.no_grad():
        want = O.qedsr_forward(x, meta, sd, res_scale=0.1)
        pol = O.qedsr_forward(x, meta, sd, res_scale=0.1, nm=O.Numerics(torch.bfloat16))
        out = net.cuda().eval()(x.cuda(), meta.cuda()).cpu()
    pol_err = max_norm_err(pol, want)
    assert out.shape == want.shape
    assert max_norm_err(out, want) <= 2.0 * pol_err + 1e-4, (max_norm_err(out, want), pol_err)


def test_net_run_and_process_device_postprocessing_is_bit_identical_to_the_host_path(tmp_path):
    """ModelInterface.net_run_and_process: clip + RGB->YCbCr run on the GPU (dfir_postprocess_rgb); both returned arrays
    must equal, bit for bit, what the reference's numpy post-processing gives on the same network output"""
    import numpy as np
    from SISR.models import ModelInterface
    torch.manual_seed(8)
    mi = ModelInterface.__new__(ModelInterface)
    mi.model = ModelInterface.define_model("qrcan", device=0, model_save_dir=str(tmp_path), eval_mode=True,
                                           metadata=["blur_kernel"], n_resgroups=1, n_resblocks=2, n_feats=64, scale=4,
                                           style="standard", include_q_layer=True, precision="fp32")
    mi.configuration = {"input": "unmodified", "colorspace": mi.model.colorspace}
    g = torch.Generator().manual_seed(5)
    lr = torch.rand(3, 3, 20, 28, generator=g) * 1.6 - 0.3          # pushes outputs outside [0,1]: the clip matters
    meta = torch.rand(3, 10, generator=g, dtype=torch.float64) * 0.4
    keys = [("blur_kernel",) * 3] * 10
    rgb, ycc, loss, timing = mi.net_run_and_process(lr=lr, hr=None, metadata=meta, metadata_keys=keys)
    raw = mi.model.run_eval(lr, metadata=meta, metadata_keys=keys)[0]
    want_rgb, want_ycc, _, _ = mi._host_postprocess(raw, None, None)
    assert isinstance(rgb, np.ndarray) and rgb.dtype == np.float32 and rgb.shape == (3, 3, 80, 112)
    assert rgb.min() >= 0.0 and rgb.max() <= 1.0
    assert np.array_equal(rgb, want_rgb)
    assert np.array_equal(ycc, want_ycc)
    # 8-bit results + Y-channel PSNR on the device: against the host expressions on the fp32 arrays
    from sr_tools.metrics import psnr
    hr = torch.rand(3, 3, 80, 112, generator=g)
    rgb8, ycc8, _, _ = mi.net_run_and_process(lr=lr, hr=hr, metadata=meta, metadata_keys=keys, output_dtype="uint8",
                                              device_psnr=True)
    assert rgb8.dtype == np.uint8 and ycc8.dtype == np.uint8 and rgb8.shape == rgb.shape
    assert np.array_equal(rgb8, np.rint(want_rgb * 255.0).astype(np.uint8))
    assert np.array_equal(ycc8, np.rint(np.clip(want_ycc, 0, 1) * 255.0).astype(np.uint8))
    hr_ycc = mi.colorspace_convert(hr, colorspace='rgb')
    for i in range(3):
        want = psnr(want_ycc[i, 0], hr_ycc[i, 0], max_value=1.0)
        assert abs(float(mi.last_y_psnr[i]) - float(want)) <= 1e-4, (i, mi.last_y_psnr[i], want)
    # identical images: the reference's convention (100 dB)
    same = torch.from_numpy(want_rgb)
    mi.net_run_and_process(lr=lr, hr=same, metadata=meta, metadata_keys=keys, device_psnr=True)
    assert mi.last_y_psnr is not None and np.all(np.isfinite(mi.last_y_psnr))


@pytest.mark.parametrize("name", ["rcan_g2b2", "edsr_f64_b3", "san_g2b2", "han_b1"])
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_non_meta_baselines_match_reference_golden(name, precision):
    """RCAN / EDSR (advanced/architectures.py) through the Q-net kernels with the meta scale == 1:
    forward(x) without metadata, fp32 mode <= 1e-4, bf16 mode within the storage-policy error"""
    ref, info = load_golden(name)
    net, x, meta = _build(info, precision)
    with torch.no_grad():
        out = net(x.cuda()).cpu()
    assert out.shape == ref.shape
    if precision == "fp32":
        assert max_norm_err(out, ref) <= 1e-4
    else:
        _, pol_err = _policy_error(info, ref)
        assert max_norm_err(out, ref) <= 2.0 * pol_err + 1e-4, (max_norm_err(out, ref), pol_err)


def test_rcan_handler_through_the_registry(tmp_path):
    """`[model] name = 'rcan'` dispatches to the B200 path: run_eval and one run_train step through the handler"""
    from SISR.models import ModelInterface
    torch.manual_seed(8)
    h = ModelInterface.define_model("edsr", device=0, model_save_dir=str(tmp_path), eval_mode=False, lr=1e-3,
                                    num_blocks=2, scale=2)
    g = torch.Generator().manual_seed(0)
    x = torch.rand(2, 3, 16, 16, generator=g)
    y = torch.nn.functional.interpolate(x, scale_factor=2, mode="bicubic", align_corners=False).clamp(0, 1)
    out, _, _ = h.run_eval(x)
    assert out.shape == (2, 3, 32, 32) and not out.is_cuda
    l0, _ = h.run_train(x, y)
    for _ in range(8):
        l1, _ = h.run_train(x, y)
    assert float(l1) < float(l0)


@pytest.mark.parametrize("shape", [(1, 1, 1), (2, 1, 7), (1, 9, 1), (1, 3, 129), (3, 2, 257), (300, 2, 5), (75, 4, 130)])
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_degenerate_and_ragged_image_shapes(shape, precision):
    """single pixels, single rows / columns (first row == last row in the pool-by-linearity statistics), widths one
    pixel past a multiple of the 128-pixel tile, and many tiny images (a CTA's row band touches three or more images, so
    both epilogue groups of conv2 evaluate attention vectors), against the oracle"""
    from deepfir_b200.qrcan import QRCAN
    torch.manual_seed(11)
    kw = dict(n_resgroups=1, n_resblocks=2, style="standard", num_metadata=10, include_q_layer=True, scale=2)
    net = QRCAN(precision=precision, **kw)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    b, h, w = shape
    g = torch.Generator().manual_seed(12)
    x = torch.rand(b, 3, h, w, generator=g)
    meta = torch.rand(b, 10, 1, 1, generator=g) * 0.4
    with torch.no_grad():
        want = O.qrcan_forward(x, meta, sd, style="standard")
        out = net.cuda().eval()(x.cuda(), meta.cuda()).cpu()
    assert out.shape == want.shape
    if precision == "fp32":
        assert max_norm_err(out, want) <= 1e-4
    else:
        pol = O.qrcan_forward(x, meta, sd, style="standard", nm=O.Numerics(torch.bfloat16))
        assert max_norm_err(out, want) <= 2.0 * max_norm_err(pol, want) + 1e-4


def test_empty_batch_returns_an_empty_result():
    from deepfir_b200.qrcan import QRCAN
    net = QRCAN(n_resgroups=1, n_resblocks=1, style="standard", num_metadata=10, include_q_layer=True, scale=4).cuda().eval()
    with torch.no_grad():
        out = net(torch.zeros(0, 3, 8, 8, device="cuda"), torch.zeros(0, 10, 1, 1, device="cuda"))
    assert out.shape == (0, 3, 32, 32)


def test_cpu_tensors_and_missing_library_fail_loudly():
    """there is no CPU / eager fallback: a CPU tensor raises instead of silently running somewhere else"""
    from deepfir_b200.qrcan import QRCAN
    net = QRCAN(n_resgroups=1, n_resblocks=1, style="standard", num_metadata=10, include_q_layer=True, scale=4)
    with pytest.raises(RuntimeError):
        net(torch.zeros(1, 3, 8, 8), torch.zeros(1, 10, 1, 1))


def test_san_handler_chops_like_the_reference(tmp_path):
    """`san` through the registry: run_eval always goes through forward_chop (4 overlapping quadrants, shave 10)"""
    from SISR.models import ModelInterface
    ref, info = load_golden("san_g2b2")
    h = ModelInterface.define_model("san", device=0, model_save_dir=str(tmp_path), eval_mode=True, precision="fp32",
                                    max_combined_im_size=600)
    x = torch.rand(1, 3, 28, 24, generator=torch.Generator().manual_seed(2))
    out, _, _ = h.run_eval(x)
    assert out.shape == (1, 3, 112, 96) and torch.isfinite(out).all()
    sd = {k: v.detach().cpu() for k, v in h.net.state_dict().items()}
    hs, ws = 14 + 10, 12 + 10
    with torch.no_grad():
        q0 = O.san_forward(x[:, :, :hs, :ws], sd)       # top-left quadrant as the reference would run it
    assert max_norm_err(out[:, :, :56, :48], q0[:, :, :56, :48]) <= 1e-4


def test_run_eval_overlaps_the_host_copy_of_half_batches_without_changing_values(tmp_path):
    """BaseModel.run_eval with `overlap_d2h_min_batch` set runs two half batches and copies the first to the host while the
    second is computed (opt-in: slower at the BASELINE shape): same values, bit for bit, as the single pass (an image never
    depends on the rest of its batch)"""
    from SISR.models import ModelInterface
    torch.manual_seed(8)
    h = ModelInterface.define_model("qrcan", device=0, model_save_dir=str(tmp_path), eval_mode=True, metadata=["blur_kernel"],
                                    n_resgroups=2, n_resblocks=2, n_feats=64, scale=2, style="standard", include_q_layer=True,
                                    precision="bf16")
    g = torch.Generator().manual_seed(1)
    x = torch.rand(16, 3, 20, 24, generator=g)
    meta = torch.rand(16, 10, generator=g, dtype=torch.float64) * 0.4
    keys = [("blur_kernel",) * 16] * 10
    assert h.overlap_d2h_min_batch is None
    b = h.run_eval(x, metadata=meta, metadata_keys=keys)[0].clone()
    h.overlap_d2h_min_batch = 16
    a = h.run_eval(x, metadata=meta, metadata_keys=keys)[0]
    assert a.shape == b.shape == (16, 3, 40, 48) and not a.is_cuda
    assert torch.equal(a, b)
