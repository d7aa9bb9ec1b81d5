"""-m gpu: the training step (forward with saved activations + backward) of the B200 path against autograd through
the CPU oracle, operator by operator and end to end through the C ABI."""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from deepfir_b200 import _lib
from tests.golden_util import (NO_TRAINING_PATH, case_tensors, load_golden, oracle_grads,
                               trainable_grad_golden_names)
from tests.gpu_util import bf16_round, lib, nhwc_bf16, nhwc_f32, stream, sync

pytestmark = pytest.mark.gpu


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def torch_wgrad(x_nchw, dy_nchw):
    """reference weight / bias gradient of a 3x3 zero-padded conv via autograd (fp64)"""
    cin, cout = x_nchw.shape[1], dy_nchw.shape[1]
    w = torch.zeros(cout, cin, 3, 3, dtype=torch.float64, requires_grad=True)
    b = torch.zeros(cout, dtype=torch.float64, requires_grad=True)
    out = F.conv2d(x_nchw.double(), w, b, padding=1)
    out.backward(dy_nchw.double())
    return w.grad.float(), b.grad.float()


@pytest.mark.parametrize("B,H,W,cin,cout", [(2, 7, 9, 64, 64), (1, 5, 40, 16, 24), (2, 6, 33, 64, 256), (1, 4, 4, 256, 64)])
def test_wgrad_f32(B, H, W, cin, cout):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, cin, H, W, generator=g)
    dy = torch.randn(B, cout, H, W, generator=g)
    dw_ref, db_ref = torch_wgrad(x, dy)
    n = lib().dfir_conv3x3_wgrad_scratch_bytes(B, H, W, cin, cout, 1)
    scratch = torch.empty(n, dtype=torch.uint8, device="cuda")
    dw = torch.full((cout, cin, 3, 3), float("nan"), device="cuda")
    db = torch.full((cout,), float("nan"), device="cuda")
    xd, dyd = nhwc_f32(x), nhwc_f32(dy)
    _lib.check(lib().dfir_conv3x3_wgrad_f32(dyd.data_ptr(), xd.data_ptr(), B, H, W, cin, cout, dw.data_ptr(), db.data_ptr(),
                                            scratch.data_ptr(), n, stream()), "wgrad f32")
    sync()
    assert rel(dw.cpu(), dw_ref) <= 1e-5
    assert rel(db.cpu(), db_ref) <= 1e-5


@pytest.mark.parametrize("B,H,W", [(1, 1, 1), (2, 7, 9), (1, 5, 128), (2, 3, 150), (3, 40, 64), (16, 64, 64)])
def test_wgrad_c64_tensor_core(B, H, W):
    """bf16 operands, exact products, fp32 accumulation: against fp64 autograd on the same bf16-rounded operands"""
    g = torch.Generator().manual_seed(2)
    x = bf16_round(torch.randn(B, 64, H, W, generator=g))
    dy = bf16_round(torch.randn(B, 64, H, W, generator=g))
    dw_ref, db_ref = torch_wgrad(x, dy)
    n = lib().dfir_conv3x3_wgrad_scratch_bytes(B, H, W, 64, 64, 0)
    scratch = torch.empty(n, dtype=torch.uint8, device="cuda")
    dw = torch.full((64, 64, 3, 3), float("nan"), device="cuda")
    db = torch.full((64,), float("nan"), device="cuda")
    xd, dyd = nhwc_bf16(x), nhwc_bf16(dy)
    _lib.check(lib().dfir_conv3x3_wgrad_c64(dyd.data_ptr(), 0, 0, 0, xd.data_ptr(), B, H, W, dw.data_ptr(), db.data_ptr(),
                                            0, 1, scratch.data_ptr(), n, stream()), "wgrad c64")
    sync()
    assert rel(dw.cpu(), dw_ref) <= 2e-5
    assert rel(db.cpu(), db_ref) <= 2e-5


def test_wgrad_c64_strided_pixel_shuffle_slice():
    """weight gradient of one sub-pixel phase of an upsampler conv: dY is a strided view of the PixelShuffle output's
    gradient, the result lands in rows s + n*r^2 of the [256][64][3][3] gradient"""
    B, h, w, r = 2, 6, 10, 2
    g = torch.Generator().manual_seed(3)
    x = bf16_round(torch.randn(B, 64, h, w, generator=g))
    dU = bf16_round(torch.randn(B, 64, h * r, w * r, generator=g))   # gradient of the shuffled output
    dy_full = F.pixel_unshuffle(dU, r)                               # [B][256][h][w], channel c*4 + i*2 + j
    dw_ref, db_ref = torch_wgrad(x, dy_full)
    n = lib().dfir_conv3x3_wgrad_scratch_bytes(B, h, w, 64, 64, 0)
    scratch = torch.empty(n, dtype=torch.uint8, device="cuda")
    dw = torch.full((256, 64, 3, 3), float("nan"), device="cuda")
    db = torch.full((256,), float("nan"), device="cuda")
    xd, dUd = nhwc_bf16(x), nhwc_bf16(dU)
    W2 = w * r
    for s in range(r * r):
        i, j = divmod(s, r)
        base = dUd.data_ptr() + (i * W2 + j) * 64 * 2
        _lib.check(lib().dfir_conv3x3_wgrad_c64(base, r * 128, r * W2 * 128, h * r * W2 * 128, xd.data_ptr(), B, h, w,
                                                dw.data_ptr(), db.data_ptr(), s, r * r, scratch.data_ptr(), n, stream()),
                   "wgrad slice")
    sync()
    assert rel(dw.cpu(), dw_ref) <= 2e-5
    assert rel(db.cpu(), db_ref) <= 2e-5


@pytest.mark.parametrize("B,H,W,mode", [(2, 7, 9, "plain"), (1, 5, 150, "skip"), (2, 9, 64, "mask"), (1, 3, 130, "mask")])
def test_dgrad_c64_tensor_core(B, H, W, mode):
    """data gradient = forward kernel on the transposed / rotated weights (+ ReLU mask | + fp32 accumulate)"""
    g = torch.Generator().manual_seed(4)
    wgt = bf16_round(torch.randn(64, 64, 3, 3, generator=g) / 24)
    dy = bf16_round(torch.randn(B, 64, H, W, generator=g))
    xin = torch.zeros(B, 64, H, W, dtype=torch.float64, requires_grad=True)
    F.conv2d(xin, wgt.double(), None, padding=1).backward(dy.double())
    want = xin.grad.float()
    wT = torch.empty(9 * 64 * 128, dtype=torch.uint8, device="cuda")
    wd = wgt.contiguous().cuda()
    _lib.check(lib().dfir_pack_conv3x3_bf16_ex(wd.data_ptr(), wT.data_ptr(), 64, 64, 0, 1, 1, stream()), "pack T")
    dyd = nhwc_bf16(dy)
    out_bf = torch.full((B, H, W, 64), float("nan"), dtype=torch.bfloat16, device="cuda")
    mask = skip = out32 = None
    if mode == "mask":
        t = bf16_round(torch.relu(torch.randn(B, 64, H, W, generator=g)))
        mask = nhwc_bf16(t)
        want = want * (t > 0)
    else:
        out32 = torch.full((B, H, W, 64), float("nan"), device="cuda")
        if mode == "skip":
            sk = torch.randn(B, 64, H, W, generator=g)
            out32.copy_(nhwc_f32(sk))  # in-place accumulate: skip aliases the output
            skip = out32
            want = want + sk
    ptr = lambda t: None if t is None else t.data_ptr()
    _lib.check(lib().dfir_conv3x3_c64_dgrad(dyd.data_ptr(), 0, 0, 0, wT.data_ptr(), ptr(mask), ptr(skip), ptr(out32),
                                            out_bf.data_ptr(), B, H, W, stream()), "dgrad")
    sync()
    if out32 is not None:
        assert rel(out32.cpu().permute(0, 3, 1, 2), want) <= 2e-5
    assert rel(out_bf.float().cpu().permute(0, 3, 1, 2), want) <= 6e-3


@pytest.mark.parametrize("tail_mode,feat_bf16", [(0, 0), (1, 0), (1, 1)])
def test_wgrad_small_head_and_tail(tail_mode, feat_bf16):
    B, H, W = 2, 9, 21
    g = torch.Generator().manual_seed(5)
    img = torch.randn(B, 3, H, W, generator=g)
    feat = torch.randn(B, 64, H, W, generator=g)
    if feat_bf16:
        feat = bf16_round(feat)
    if tail_mode:   # conv 64 -> 3: x = feat, dy = img
        dw_ref, db_ref = torch_wgrad(feat, img)
    else:           # conv 3 -> 64: x = img, dy = feat
        dw_ref, db_ref = torch_wgrad(img, feat)
    n = lib().dfir_conv3x3_wgrad_small_scratch_bytes(B, H, 64)
    scratch = torch.empty(n, dtype=torch.uint8, device="cuda")
    dw = torch.full(tuple(dw_ref.shape), float("nan"), device="cuda")
    db = torch.full(tuple(db_ref.shape), float("nan"), device="cuda")
    fd = nhwc_bf16(feat) if feat_bf16 else nhwc_f32(feat)
    imd = img.contiguous().cuda()
    _lib.check(lib().dfir_conv3x3_wgrad_small(imd.data_ptr(), fd.data_ptr(), feat_bf16, B, H, W, 64, 3, tail_mode,
                                              dw.data_ptr(), db.data_ptr(), scratch.data_ptr(), n, stream()), "wgrad small")
    sync()
    assert rel(dw.cpu(), dw_ref) <= 1e-5
    assert rel(db.cpu(), db_ref) <= 1e-5


# ------------------------------------------------------------------------------------------------ whole step
def _build(info, precision):
    from deepfir_b200.baselines import EDSR, HAN, RCAN, SAN
    from deepfir_b200.han_san import QHAN, QSAN
    from deepfir_b200.qrcan import QEDSR, QRCAN
    cls = {"qedsr": QEDSR, "qrcan": QRCAN, "rcan": RCAN, "edsr": EDSR, "qsan": QSAN, "qhan": QHAN, "san": SAN,
           "han": HAN}[info["model"]]
    net = cls(precision=precision, **info["kwargs"])
    sd, x, meta = case_tensors(info)
    net.load_state_dict(sd, strict=True)
    return net.cuda().train(), sd, x, meta


def _step_grads(net, x, meta, y):
    net.zero_grad(set_to_none=True)
    out = net(x.cuda(), meta.cuda())
    loss = F.l1_loss(out, y.cuda())
    loss.backward()
    # (parameters a network owns but never calls -- Q-SAN's `conv_last`, `RG.*.gamma`, `non_local.soca` -- keep .grad None,
    # as under the reference's autograd)
    return float(loss.detach()), out.detach().cpu(), {k: p.grad.detach().cpu().clone() for k, p in net.named_parameters()
                                                      if p.grad is not None}


@pytest.mark.parametrize("name", trainable_grad_golden_names())
def test_train_step_fp32_gradients_match_oracle(name):
    """per-parameter ||g - g_ref|| / ||g_ref|| <= 1e-3 in fp32 mode (g_ref = autograd through the oracle,
    itself pinned to the reference's loss.backward() by tests/test_train_cpu.py)"""
    _, info = load_golden(name)
    net, sd, x, meta = _build(info, "fp32")
    ref_loss, ref_out, y, ref = oracle_grads(info, sd, x, meta)
    loss, out, grads = _step_grads(net, x, meta, y)
    assert abs(loss - ref_loss) <= 1e-5 * abs(ref_loss)
    gmax = max(float(v.norm()) for v in ref.values())
    assert set(grads) == set(ref)
    for k, g in grads.items():
        err = float((g.double() - ref[k].double()).norm())
        assert err <= 1e-3 * float(ref[k].norm()) + 1e-6 * gmax, (k, err, float(ref[k].norm()))


@pytest.mark.parametrize("name", [n for n in trainable_grad_golden_names() if "f256" not in n])
def test_train_step_bf16_gradients_close_to_oracle(name):
    """tensor-core path: bf16 operands (activations, weights and gradients), fp32 accumulation and fp32 gradient
    stream.  Bar: loss within 1e-3 relative, every gradient within 4e-2 of its norm (+ a floor of 1e-3 of the largest
    gradient norm for the near-zero ones)"""
    _, info = load_golden(name)
    net, sd, x, meta = _build(info, "bf16")
    ref_loss, ref_out, y, ref = oracle_grads(info, sd, x, meta)
    loss, out, grads = _step_grads(net, x, meta, y)
    assert abs(loss - ref_loss) <= 1e-3 * abs(ref_loss)
    gmax = max(float(v.norm()) for v in ref.values())
    worst = 0.0
    for k, g in grads.items():
        err = float((g.double() - ref[k].double()).norm())
        worst = max(worst, err / max(float(ref[k].norm()), 1e-3 * gmax))
        assert err <= 4e-2 * float(ref[k].norm()) + 1e-3 * gmax, (k, err, float(ref[k].norm()))
    print("bf16 worst relative gradient error", worst)


def test_train_forward_schedules_agree(monkeypatch):
    """the training forward has two block schedules (streamer for short bands, pool-by-linearity for long ones): same
    loss and gradients up to bf16 rounding of r"""
    _, info = load_golden("qrcan_standard_g2b2")
    res = {}
    for sched in ("streamer", "linear"):
        monkeypatch.setenv("DFIR_TRAIN_SCHEDULE", sched)
        net, sd, x, meta = _build(info, "bf16")
        net.cuda_graphs = False
        _, _, y, _ = oracle_grads(info, sd, x, meta)
        res[sched] = _step_grads(net, x, meta, y)
    assert abs(res["linear"][0] - res["streamer"][0]) <= 1e-3 * abs(res["streamer"][0])
    gmax = max(float(v.norm()) for v in res["streamer"][2].values())
    for k, g in res["linear"][2].items():
        ref = res["streamer"][2][k]
        assert float((g - ref).norm()) <= 3e-2 * float(ref.norm()) + 1e-3 * gmax, k


def test_gradients_accumulate_without_zero_grad():
    _, info = load_golden("qrcan_noq_scale2")
    net, sd, x, meta = _build(info, "fp32")
    _, _, y, _ = oracle_grads(info, sd, x, meta)
    _, _, g1 = _step_grads(net, x, meta, y)
    out = net(x.cuda(), meta.cuda())          # second backward on top of the existing .grad
    F.l1_loss(out, y.cuda()).backward()
    for k, p in net.named_parameters():
        assert rel(p.grad.cpu(), 2 * g1[k]) <= 1e-5, k


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_optimizer_step_then_inference_uses_fresh_weights(precision):
    """after optimizer.step() the kernel-format weights are refreshed by dfir_qrcan_repack: the inference forward must
    equal the oracle on the UPDATED state_dict"""
    from tests.golden_util import max_norm_err, oracle_forward
    _, info = load_golden("qrcan_standard_g2b2")
    net, sd, x, meta = _build(info, precision)
    _, _, y, _ = oracle_grads(info, sd, x, meta)
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    for _ in range(2):
        opt.zero_grad()
        F.l1_loss(net(x.cuda(), meta.cuda()), y.cuda()).backward()
        opt.step()
    new_sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
    with torch.no_grad():
        out = net(x.cuda(), meta.cuda()).cpu()
        want = oracle_forward(info, new_sd, x, meta)
    assert max_norm_err(out, want) <= (1e-4 if precision == "fp32" else 3e-2)
    assert max_norm_err(want, oracle_forward(info, sd, x, meta)) > 1e-3  # the step really moved the weights


def test_handler_run_train_reduces_the_loss(tmp_path):
    """BaseModel.run_train through the handler registry (models/__init__.py:466-479): L1 + Adam + scheduler"""
    from SISR.models import ModelInterface
    torch.manual_seed(8)
    h = ModelInterface.define_model("qrcan", device=0, model_save_dir=str(tmp_path), eval_mode=False, lr=1e-3,
                                    metadata=["blur_kernel"], n_resgroups=2, n_resblocks=2, n_feats=64, scale=2,
                                    style="standard", include_q_layer=True, precision="bf16",
                                    scheduler="cosine_annealing_warm_restarts",
                                    scheduler_params=dict(t_mult=1, restart_period=100, lr_min=1e-6))
    g = torch.Generator().manual_seed(0)
    x = torch.rand(4, 3, 16, 16, generator=g)
    y = F.interpolate(x, scale_factor=2, mode="bicubic", align_corners=False).clamp(0, 1)
    meta = torch.rand(4, 10, generator=g, dtype=torch.float64) * 0.4
    keys = [("blur_kernel",) * 4] * 10
    losses = []
    for _ in range(12):
        loss, out = h.run_train(x, y, metadata=meta, metadata_keys=keys)
        assert out.shape == y.shape and not out.is_cuda
        losses.append(float(loss))
    assert all(np.isfinite(losses))
    assert losses[-1] < 0.7 * losses[0], losses


def test_training_with_selective_meta_blocks_and_changing_batch_shapes():
    """q layers only in some groups / blocks (selective_meta_blocks, num_q_layers_inner_residual), and two batch shapes
    alternating through the per-shape CUDA graphs: gradients against autograd through the oracle each time"""
    from deepfir_b200.qrcan import QRCAN
    from oracle import deepfir_oracle as O
    torch.manual_seed(21)
    kw = dict(n_resgroups=3, n_resblocks=3, style="max_concat", num_metadata=10, include_q_layer=True, scale=2,
              selective_meta_blocks=[True, False, True], num_q_layers_inner_residual=2)
    net = QRCAN(precision="fp32", **kw).cuda().train()
    sd = {k: v.detach().cpu().clone() for k, v in net.state_dict().items()}
    assert sum(".q_node." in k for k in sd) == 2 * 2 * 4      # q layers in groups 0 and 2, blocks 0 and 1 only
    g = torch.Generator().manual_seed(22)
    batches = []
    for b, h, w in [(2, 10, 12), (3, 8, 8)]:
        batches.append((torch.rand(b, 3, h, w, generator=g), torch.rand(b, 10, 1, 1, generator=g) * 0.4,
                        torch.rand(b, 3, 2 * h, 2 * w, generator=g)))
    for rep in range(3):                                      # eager, graph capture, graph replay
        for x, meta, y in batches:
            leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
            F.l1_loss(O.qrcan_forward(x, meta, leaves, style="max_concat"), y).backward()
            net.zero_grad(set_to_none=True)
            F.l1_loss(net(x.cuda(), meta.cuda()), y.cuda()).backward()
            gmax = max(float(v.grad.norm()) for v in leaves.values())
            for k, p in net.named_parameters():
                err = float((p.grad.cpu().double() - leaves[k].grad.double()).norm())
                assert err <= 1e-3 * float(leaves[k].grad.norm()) + 1e-6 * gmax, (rep, k, err)


def test_wide_qedsr_trains_on_the_fp32_kernels_and_infers_on_the_tensor_cores(tmp_path):
    """a 128-feature Q-EDSR built with the default precision: run_train uses the library's fp32 kernels (the tensor-core
    training kernels are 64-feature), run_eval the 64-channel-plane tensor-core path; the loss goes down and the two
    paths see the same (updated) weights"""
    from SISR.models import ModelInterface
    from oracle import deepfir_oracle as O
    torch.manual_seed(8)
    h = ModelInterface.define_model("qedsr", device=0, model_save_dir=str(tmp_path), eval_mode=False, lr=1e-3,
                                    metadata=["blur_kernel"], num_features=128, num_blocks=2, scale=2)
    assert h.net.precision == "bf16"
    g = torch.Generator().manual_seed(0)
    x = torch.rand(2, 3, 16, 16, generator=g)
    y = F.interpolate(x, scale_factor=2, mode="bicubic", align_corners=False).clamp(0, 1)
    meta = torch.rand(2, 10, generator=g, dtype=torch.float64) * 0.4
    keys = [("blur_kernel",) * 2] * 10
    losses = [float(h.run_train(x, y, metadata=meta, metadata_keys=keys)[0]) for _ in range(8)]
    assert losses[-1] < losses[0]
    out = h.run_eval(x, metadata=meta, metadata_keys=keys)[0]
    sd = {k: v.detach().cpu() for k, v in h.net.state_dict().items()}
    with torch.no_grad():
        want = O.qedsr_forward(x, meta.float().reshape(2, 10, 1, 1), sd, res_scale=0.1)
    assert float((out - want).abs().max() / want.abs().max()) <= 3e-2


def test_flat_adam_matches_torch_adam_and_checkpoints_like_it(tmp_path):
    """FlatAdam (one kernel over flat buffers) against torch.optim.Adam on the same network, data and steps; the
    optimizer state_dict keeps torch's layout and survives a save / load into a fresh handler"""
    from deepfir_b200.flat_adam import FlatAdam
    from deepfir_b200.qrcan import QRCAN
    kw = dict(n_resgroups=1, n_resblocks=2, style="standard", num_metadata=10, include_q_layer=True, scale=2)
    g = torch.Generator().manual_seed(5)
    x = torch.rand(2, 3, 12, 12, generator=g).cuda(); y = torch.rand(2, 3, 24, 24, generator=g).cuda()
    meta = (torch.rand(2, 10, 1, 1, generator=g) * 0.4).cuda()
    nets, opts = [], []
    for cls in (torch.optim.Adam, FlatAdam):
        torch.manual_seed(9)
        net = QRCAN(precision="fp32", **kw).cuda().train()
        nets.append(net)
        opts.append(cls(net.parameters(), lr=1e-3, betas=(0.9, 0.99)))

    def step(net, opt):
        opt.zero_grad()
        F.l1_loss(net(x, meta), y).backward()
        opt.step()

    for _ in range(4):
        for net, opt in zip(nets, opts):
            step(net, opt)
    for (k, a), (_, b) in zip(nets[0].named_parameters(), nets[1].named_parameters()):
        assert float((a - b).detach().abs().max()) <= 2e-6 + 1e-5 * float(a.detach().abs().max()), k
    # inference after the flat steps sees the updated weights
    with torch.no_grad():
        nets[0].eval(); nets[1].eval()
        assert float((nets[0](x, meta) - nets[1](x, meta)).abs().max()) <= 1e-4
        nets[0].train(); nets[1].train()
    # checkpoint: torch's state_dict layout, interchangeable between the two optimizers
    sd_flat, sd_ref = opts[1].state_dict(), opts[0].state_dict()
    assert set(sd_flat["state"][0]) == set(sd_ref["state"][0]) == {"step", "exp_avg", "exp_avg_sq"}
    assert float(sd_flat["state"][0]["step"]) == float(sd_ref["state"][0]["step"]) == 4.0
    torch.save(sd_flat, tmp_path / "opt.pt")
    # a STOCK Adam resumes from the checkpoint: every parameter has its own step tensor (the flat optimizer shares one
    # internally; a checkpoint that kept the aliasing would advance the step once per parameter and iteration)
    loaded = torch.load(tmp_path / "opt.pt")
    steps = [st["step"] for st in loaded["state"].values()]
    assert len({t.untyped_storage().data_ptr() for t in steps}) == len(steps)
    torch.manual_seed(9)
    net4 = QRCAN(precision="fp32", **kw).cuda().train()
    net4.load_state_dict(nets[1].state_dict())
    opt4 = torch.optim.Adam(net4.parameters(), lr=1e-3, betas=(0.9, 0.99))
    opt4.load_state_dict(loaded)
    step(net4, opt4)
    assert all(float(st["step"]) == 5.0 for st in opt4.state.values())
    torch.manual_seed(9)
    net3 = QRCAN(precision="fp32", **kw).cuda().train()
    net3.load_state_dict(nets[1].state_dict())
    opt3 = FlatAdam(net3.parameters(), lr=1e-3, betas=(0.9, 0.99))
    opt3.load_state_dict(torch.load(tmp_path / "opt.pt"))
    step(net3, opt3); step(nets[0], opts[0])
    for (k, a), (_, b) in zip(nets[0].named_parameters(), net3.named_parameters()):
        assert float((a - b).detach().abs().max()) <= 3e-6 + 1e-5 * float(a.detach().abs().max()), k


def test_full_depth_bf16_gradients_stay_aligned_with_the_reference():
    """the published depth (10 groups x 20 RCABs): bf16 operands through 400+ layers forward and backward.  The gradient
    of every parameter tensor must point where the fp32 reference gradient points (cosine) and have its size; the fp32
    mode stays within the 1e-3 bar at this depth too."""
    _, info = load_golden("qrcan_standard_full")
    ref_loss, _, y, ref = None, None, None, None
    for precision, cos_min, rel_max in (("fp32", 0.999999, 1e-3), ("bf16", 0.999, 5e-2)):
        net, sd, x, meta = _build(info, precision)
        if ref is None:
            ref_loss, _, y, ref = oracle_grads(info, sd, x, meta)
        loss, _, grads = _step_grads(net, x, meta, y)
        assert abs(loss - ref_loss) <= (1e-5 if precision == "fp32" else 2e-3) * abs(ref_loss)
        gmax = max(float(v.norm()) for v in ref.values())
        worst_cos, worst_rel = 1.0, 0.0
        for k, g in grads.items():
            r = ref[k].double().reshape(-1)
            gd = g.double().reshape(-1)
            if float(r.norm()) < 1e-3 * gmax:
                continue  # near-zero gradients (biases deep in the trunk) carry no direction worth comparing
            cos = float(torch.dot(gd, r) / (gd.norm() * r.norm()))
            rel = float((gd - r).norm() / r.norm())
            worst_cos, worst_rel = min(worst_cos, cos), max(worst_rel, rel)
        print("%s full depth: worst cosine %.6f, worst relative error %.3e" % (precision, worst_cos, worst_rel))
        assert worst_cos >= cos_min and worst_rel <= rel_max, (precision, worst_cos, worst_rel)


def test_training_a_network_without_backward_kernels_fails_loudly():
    """every golden network now trains (NO_TRAINING_PATH is empty); what is left without a training path is a
    192-feature Q-EDSR (bf16 inference only: the fp32 training kernels need a divisor of 256): a training step must
    raise NotImplementedError, not fall back to anything"""
    from deepfir_b200.qrcan import QEDSR
    assert NO_TRAINING_PATH == ()
    net = QEDSR(num_features=192, num_blocks=1, input_para=10, scale=2).cuda().train()
    x = torch.rand(1, 3, 8, 8).cuda()
    meta = torch.rand(1, 10, 1, 1).cuda()
    with pytest.raises(NotImplementedError):
        out = net(x, meta)
        F.l1_loss(out, torch.zeros_like(out)).backward()


def test_inputs_that_require_grad_are_refused():
    """the library's backward produces parameter gradients only: an input that asks for a gradient must raise, not get
    a silent None"""
    _, info = load_golden("qrcan_noq_scale2")
    net, sd, x, meta = _build(info, "fp32")
    xg = x.cuda().requires_grad_(True)
    with pytest.raises(RuntimeError, match="gradients with respect to the input"):
        net(xg, meta.cuda())


def test_invalidate_packed_after_a_write_through_data():
    """`p.data.mul_()` does not bump Parameter._version, so the packed kernel weights would go stale: the documented
    remedy is net.invalidate_packed() (load_state_dict and train() -> eval() call it themselves)"""
    _, info = load_golden("qrcan_noq_scale2")
    for precision in ("fp32", "bf16"):
        net, sd, x, meta = _build(info, precision)
        net.eval()
        with torch.no_grad():
            a = net(x.cuda(), meta.cuda()).clone()
            tail = [p for k, p in net.named_parameters() if k.endswith("tail.1.weight")][0]
            v = tail._version
            tail.data.mul_(2.0)
            assert tail._version == v          # the write is invisible to the version counter
            net.invalidate_packed()
            b = net(x.cuda(), meta.cuda())
            bias = [p for k, p in net.named_parameters() if k.endswith("tail.1.bias")][0]
            want = 2.0 * (a - bias.reshape(1, -1, 1, 1)) + bias.reshape(1, -1, 1, 1)
        assert float((b - want).abs().max()) <= 2e-2 * float(want.abs().max()) + 1e-6, precision
        # train() -> eval() transition repacks as well
        with torch.no_grad():
            tail.data.mul_(0.5)
            net.train(); net.eval()
            c = net(x.cuda(), meta.cuda())
        assert float((c - a).abs().max()) <= 1e-5 + 1e-5 * float(a.abs().max()), precision


@pytest.mark.parametrize("model", ["qsan", "qhan"])
def test_handler_run_train_of_qsan_and_qhan_reduces_the_loss(tmp_path, model):
    """BaseModel.run_train through the registry for the staged networks (models/__init__.py:466-479) at the handlers'
    published depth (Q-SAN 20 x 10, Q-HAN 10 x 20): L1 + Adam; Q-SAN's unused parameters keep .grad None (as under the
    reference's autograd), so its Adam skips them"""
    from SISR.models import ModelInterface
    torch.manual_seed(8)
    h = ModelInterface.define_model(model, device=0, model_save_dir=str(tmp_path), eval_mode=False, lr=2e-4,
                                    metadata=["blur_kernel"], precision="bf16", scale=2)
    g = torch.Generator().manual_seed(0)
    x = torch.rand(2, 3, 16, 16, generator=g)
    y = F.interpolate(x, scale_factor=2, mode="bicubic", align_corners=False).clamp(0, 1)
    meta = torch.rand(2, 10, generator=g, dtype=torch.float64) * 0.4
    keys = [("blur_kernel",) * 2] * 10
    losses = []
    for _ in range(10):
        loss, out = h.run_train(x, y, metadata=meta, metadata_keys=keys)
        assert out.shape == y.shape
        losses.append(float(loss))
    assert all(np.isfinite(losses)) and min(losses[-3:]) < losses[0], losses
    if model == "qsan":
        assert h.net.conv_last.weight.grad is None and h.net.RG[0].conv_last.weight.grad is not None
