"""-m gpu: operator-level parity of the Q-SAN / Q-HAN specific kernels (csrc/san_han.cu) through the C ABI against the CPU
oracle: covariance pooling and the Newton-Schulz square root with the reference's hand-written backward passes
(advanced/mpncov.py), SOCA including the centre crop at >= 1000 pixels (advanced/SAN_blocks.py:261-300), region non-local
attention (:104-148, 314-336), LAM and CSAM (advanced/HAN_blocks.py:24-76)."""
import numpy as np
import pytest
import torch

from oracle import deepfir_oracle as O
from oracle import np_ops
from tests import gpu_util as G

pytestmark = pytest.mark.gpu


def _scratch(nbytes):
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device="cuda")


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("shape,crop", [((2, 9, 13), 0), ((1, 128, 128), 0), ((3, 1, 7), 0), ((1, 1003, 5), 1), ((1, 6, 1002), 1)])
def test_covpool_forward_and_backward(shape, crop):
    B, H, W = shape
    g = torch.Generator().manual_seed(H * 7 + W)
    x = torch.randn(B, 64, H, W, generator=g) * 0.7 + 0.2
    xs = x
    if crop and H > 1000:
        xs = x[:, :, (H - 1000) // 2:(H - 1000) // 2 + 1000, :]
    if crop and W > 1000:
        xs = x[:, :, :, (W - 1000) // 2:(W - 1000) // 2 + 1000]
    xr = x.clone().double().requires_grad_(True)
    xsr = xr
    if crop and H > 1000:
        xsr = xr[:, :, (H - 1000) // 2:(H - 1000) // 2 + 1000, :]
    if crop and W > 1000:
        xsr = xr[:, :, :, (W - 1000) // 2:(W - 1000) // 2 + 1000]
    want = O.covpool(xsr)
    if H * W <= 200:  # the literal restatement with the MxM centring matrix
        assert np.allclose(want.detach().numpy(), np_ops.covpool_np(xs.numpy()), rtol=1e-9, atol=1e-12)
    gout = torch.randn(B, 64, 64, generator=g)
    want.backward(gout.double())
    L = G.lib()
    xd = G.nhwc_f32(x)
    cov = torch.full((B, 64, 64), float("nan"), device="cuda")
    sc = _scratch(L.dfir_covpool_scratch_bytes(B))
    assert L.dfir_covpool(xd.data_ptr(), cov.data_ptr(), sc.data_ptr(), sc.numel(), B, H, W, 64, crop, G.stream()) == 0
    gx = torch.full((B, H, W, 64), float("nan"), device="cuda")
    gd = gout.cuda().contiguous()
    assert L.dfir_covpool_backward(xd.data_ptr(), gd.data_ptr(), gx.data_ptr(), sc.data_ptr(), sc.numel(), B, H, W, 64, crop,
                                   G.stream()) == 0
    G.sync()
    assert _rel(cov.cpu(), want.detach()) <= 2e-5
    assert _rel(G.to_nchw(gx), xr.grad) <= 2e-5
    # the reference's own formula (Covpool.backward, mpncov.py:35-47) on the small cases
    if H * W <= 200:
        M = H * W
        ihat = torch.full((M, M), -1.0 / M / M, dtype=torch.float64)
        ihat.diagonal().add_(1.0 / M)
        ref = (gout.double() + gout.double().transpose(1, 2)).bmm(x.double().reshape(B, 64, M)).matmul(ihat)
        assert _rel(G.to_nchw(gx).reshape(B, 64, M), ref) <= 2e-5


@pytest.mark.parametrize("iters", [5, 2, 3])
def test_sqrtm_forward_and_backward(iters):
    """Newton-Schulz square root and the reference's closed-form backward (Sqrtm.backward, mpncov.py:78-112) against
    autograd through the oracle's forward in float64"""
    B = 3
    g = torch.Generator().manual_seed(11 + iters)
    x = torch.randn(B, 64, 9, 14, generator=g)
    cov = O.covpool(x.double()).float() + 1e-3 * torch.eye(64)
    cr = cov.clone().double().requires_grad_(True)
    want = O.sqrtm_ns(cr, iters)
    gout = torch.randn(B, 64, 64, generator=g)
    want.backward(gout.double())
    L = G.lib()
    cd, gd = cov.cuda().contiguous(), gout.cuda().contiguous()
    out = torch.full((B, 64, 64), float("nan"), device="cuda")
    gin = torch.full((B, 64, 64), float("nan"), device="cuda")
    assert L.dfir_sqrtm(cd.data_ptr(), out.data_ptr(), B, 64, iters, G.stream()) == 0
    sc = _scratch(L.dfir_sqrtm_scratch_bytes(B, iters))
    assert L.dfir_sqrtm_backward(cd.data_ptr(), gd.data_ptr(), gin.data_ptr(), sc.data_ptr(), sc.numel(), B, 64, iters,
                                 G.stream()) == 0
    G.sync()
    assert _rel(out.cpu(), want.detach()) <= 5e-5
    assert _rel(gin.cpu(), cr.grad) <= 2e-4


@pytest.mark.parametrize("shape", [(2, 12, 10), (1, 74, 74), (1, 1003, 4), (1, 3, 1001)])
def test_soca_vector_including_the_1000_pixel_centre_crop(shape):
    B, H, W = shape
    g = torch.Generator().manual_seed(H + 3 * W)
    x = torch.randn(B, 64, H, W, generator=g) * 0.5
    R = 4
    sd = {"s.conv_du.0.weight": torch.randn(R, 64, 1, 1, generator=g) * 0.3, "s.conv_du.0.bias": torch.randn(R, generator=g) * 0.1,
          "s.conv_du.2.weight": torch.randn(64, R, 1, 1, generator=g) * 0.3, "s.conv_du.2.bias": torch.randn(64, generator=g) * 0.1}
    want = O.soca_vector(x, sd, "s").reshape(B, 64)
    mlp = torch.cat([sd["s.conv_du.0.weight"].reshape(-1), sd["s.conv_du.0.bias"], sd["s.conv_du.2.weight"].reshape(-1),
                     sd["s.conv_du.2.bias"]]).cuda()
    L = G.lib()
    xd = G.nhwc_f32(x)
    sv = torch.full((B, 64), float("nan"), device="cuda")
    sc = _scratch(L.dfir_soca_scratch_bytes(B))
    assert L.dfir_soca(xd.data_ptr(), mlp.data_ptr(), R, sv.data_ptr(), sc.data_ptr(), B, H, W, 64, G.stream()) == 0
    G.sync()
    assert float((sv.cpu() - want).abs().max()) <= 2e-5


@pytest.mark.parametrize("shape", [(2, 8, 12), (1, 9, 7), (1, 26, 30), (2, 4, 5)])
def test_region_nonlocal_attention(shape):
    """Nonlocal_CA: 2x2 regions (odd sizes give unequal regions), theta/phi/g 1x1 convs to 8 channels, always-on 2x2
    max-pool of phi and g, softmax attention, W conv + x"""
    B, H, W = shape
    g = torch.Generator().manual_seed(H * 5 + W)
    x = torch.randn(B, 64, H, W, generator=g)
    p = "n.non_local"
    sd = {}
    for name, (co, ci) in {"theta": (8, 64), "phi.0": (8, 64), "g.0": (8, 64), "W": (64, 8)}.items():
        sd["%s.%s.weight" % (p, name)] = torch.randn(co, ci, 1, 1, generator=g) * 0.2
        sd["%s.%s.bias" % (p, name)] = torch.randn(co, generator=g) * 0.1
    want = O.nonlocal_ca(x, sd, "n")
    w_tpg = torch.cat([sd[p + ".theta.weight"].reshape(8, 64), sd[p + ".phi.0.weight"].reshape(8, 64),
                       sd[p + ".g.0.weight"].reshape(8, 64)]).contiguous().cuda()
    b_tpg = torch.cat([sd[p + ".theta.bias"], sd[p + ".phi.0.bias"], sd[p + ".g.0.bias"]]).contiguous().cuda()
    w_out, b_out = sd[p + ".W.weight"].reshape(64, 8).contiguous().cuda(), sd[p + ".W.bias"].cuda()
    L = G.lib()
    xd = G.nhwc_f32(x)
    out = torch.full((B, H, W, 64), float("nan"), device="cuda")
    sc = _scratch(L.dfir_nonlocal_scratch_bytes(B, H, W))
    assert L.dfir_nonlocal(xd.data_ptr(), w_tpg.data_ptr(), b_tpg.data_ptr(), w_out.data_ptr(), b_out.data_ptr(), out.data_ptr(),
                           sc.data_ptr(), B, H, W, 64, G.stream()) == 0
    G.sync()
    assert _rel(G.to_nchw(out), want) <= 2e-5


@pytest.mark.parametrize("N,shape", [(11, (2, 8, 8)), (3, (1, 5, 9)), (1, (1, 4, 4)), (11, (1, 32, 32))])
def test_layer_attention(N, shape):
    """LAM_Module: N x N Gram matrix over (channel, pixel), softmax(max - E), gamma * att X + X; maps read through a
    (possibly negative) stride like the Q-HAN forward does"""
    B, H, W = shape
    g = torch.Generator().manual_seed(N + H)
    x5 = torch.randn(B, N, 64, H, W, generator=g) * 0.3
    gamma = torch.tensor(0.37)
    want = O.lam(x5, gamma)                                         # [B][N*64][H][W]
    stack = torch.stack([G.nhwc_f32(x5[:, n]) for n in range(N)])   # [N][B][H][W][64]
    rev = torch.flip(stack, dims=[0]).contiguous()                  # map n lives at rev[N-1-n]: negative stride
    out = torch.full((B, H, W, N * 64), float("nan"), device="cuda")
    L = G.lib()
    sc = _scratch(L.dfir_lam_scratch_bytes(B, N))
    per_map = B * H * W * 64
    assert L.dfir_lam(rev[N - 1].data_ptr(), -per_map, float(gamma), out.data_ptr(), sc.data_ptr(), N, B, H * W, 64,
                      G.stream()) == 0
    G.sync()
    assert _rel(G.to_nchw(out), want) <= 5e-5


@pytest.mark.parametrize("shape", [(2, 6, 9), (1, 1, 1), (1, 17, 5)])
def test_channel_spatial_attention(shape):
    """CSAM_Module: 3x3x3 conv over the (channel, y, x) volume, sigmoid, x * (gamma * s) + x"""
    B, H, W = shape
    g = torch.Generator().manual_seed(H * 3 + W)
    x = torch.randn(B, 64, H, W, generator=g)
    sd = {"c.conv.weight": torch.randn(1, 1, 3, 3, 3, generator=g) * 0.3, "c.conv.bias": torch.randn(1, generator=g) * 0.1,
          "c.gamma": torch.tensor(0.6)}
    want = O.csam(x, sd, "c")
    L = G.lib()
    xd = G.nhwc_f32(x)
    w27 = sd["c.conv.weight"].reshape(-1).contiguous().cuda()
    out = torch.full((B, H, W, 64), float("nan"), device="cuda")
    assert L.dfir_csam(xd.data_ptr(), w27.data_ptr(), float(sd["c.conv.bias"]), float(sd["c.gamma"]), out.data_ptr(), B, H, W,
                       64, G.stream()) == 0
    G.sync()
    assert _rel(G.to_nchw(out), want) <= 2e-5


# ------------------------------------------------------------------------------------------------ backward operators
def _relg(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


@pytest.mark.parametrize("shape", [(2, 8, 12), (1, 9, 7), (1, 26, 30), (2, 4, 5), (1, 32, 32)])
def test_region_nonlocal_attention_backward(shape):
    """dfir_nonlocal_backward against fp64 autograd through the oracle's Nonlocal_CA (input, theta/phi/g, W gradients),
    odd regions included (pixels no pooling window covers get no phi / g gradient); a second call accumulates"""
    B, H, W = shape
    g = torch.Generator().manual_seed(H * 5 + W + 1)
    x = torch.randn(B, 64, H, W, generator=g)
    p = "n.non_local"
    sd = {}
    for name, (co, ci) in {"theta": (8, 64), "phi.0": (8, 64), "g.0": (8, 64), "W": (64, 8)}.items():
        sd["%s.%s.weight" % (p, name)] = torch.randn(co, ci, 1, 1, generator=g) * 0.2
        sd["%s.%s.bias" % (p, name)] = torch.randn(co, generator=g) * 0.1
    leaves = {k: v.double().requires_grad_(True) for k, v in sd.items()}
    xr = x.double().requires_grad_(True)
    gout = torch.randn(B, 64, H, W, generator=g)
    O.nonlocal_ca(xr, leaves, "n").backward(gout.double())
    w_tpg = torch.cat([sd[p + ".theta.weight"].reshape(8, 64), sd[p + ".phi.0.weight"].reshape(8, 64),
                       sd[p + ".g.0.weight"].reshape(8, 64)]).contiguous().cuda()
    b_tpg = torch.cat([sd[p + ".theta.bias"], sd[p + ".phi.0.bias"], sd[p + ".g.0.bias"]]).contiguous().cuda()
    w_out = sd[p + ".W.weight"].reshape(64, 8).contiguous().cuda()
    want_w = torch.cat([leaves[p + ".theta.weight"].grad.reshape(8, 64), leaves[p + ".phi.0.weight"].grad.reshape(8, 64),
                        leaves[p + ".g.0.weight"].grad.reshape(8, 64)])
    want_b = torch.cat([leaves[p + ".theta.bias"].grad, leaves[p + ".phi.0.bias"].grad, leaves[p + ".g.0.bias"].grad])
    L = G.lib()
    xd, gd = G.nhwc_f32(x), G.nhwc_f32(gout)
    nan = lambda *s: torch.full(s, float("nan"), device="cuda")
    dx, gw, gb, gwo, gbo = nan(B, H, W, 64), nan(24, 64), nan(24), nan(64, 8), nan(64)
    sc = _scratch(L.dfir_nonlocal_backward_scratch_bytes(B, H, W))
    for rep in range(2):
        assert L.dfir_nonlocal_backward(xd.data_ptr(), gd.data_ptr(), w_tpg.data_ptr(), b_tpg.data_ptr(), w_out.data_ptr(),
                                        dx.data_ptr(), gw.data_ptr(), gb.data_ptr(), gwo.data_ptr(), gbo.data_ptr(), rep,
                                        sc.data_ptr(), sc.numel(), B, H, W, 64, G.stream()) == 0
        G.sync()
        k = rep + 1
        assert _relg(G.to_nchw(dx), xr.grad) <= 2e-5
        assert _relg(gw.cpu(), k * want_w) <= 2e-5 and _relg(gb.cpu(), k * want_b) <= 2e-5
        assert _relg(gwo.cpu(), k * leaves[p + ".W.weight"].grad.reshape(64, 8)) <= 2e-5
        assert _relg(gbo.cpu(), k * leaves[p + ".W.bias"].grad) <= 2e-5


@pytest.mark.parametrize("N,shape", [(11, (2, 8, 8)), (3, (1, 5, 9)), (1, (1, 4, 4)), (11, (1, 32, 32))])
def test_layer_attention_backward(N, shape):
    B, H, W = shape
    g = torch.Generator().manual_seed(N + H + 3)
    x5 = torch.randn(B, N, 64, H, W, generator=g) * 0.3
    gamma = torch.tensor(0.37)
    xr = x5.double().requires_grad_(True)
    gr = gamma.double().requires_grad_(True)
    gout = torch.randn(B, N * 64, H, W, generator=g)
    O.lam(xr, gr).backward(gout.double())
    stack = torch.stack([G.nhwc_f32(x5[:, n]) for n in range(N)])
    rev = torch.flip(stack, dims=[0]).contiguous()
    out = torch.empty(B, H, W, N * 64, device="cuda")
    L = G.lib()
    fsc = _scratch(L.dfir_lam_scratch_bytes(B, N))
    per_map = B * H * W * 64
    assert L.dfir_lam(rev[N - 1].data_ptr(), -per_map, float(gamma), out.data_ptr(), fsc.data_ptr(), N, B, H * W, 64,
                      G.stream()) == 0
    gd = G.nhwc_f32(gout)
    drev = torch.full_like(rev, float("nan"))
    dgamma = torch.full((1,), float("nan"), device="cuda")
    sc = _scratch(L.dfir_lam_backward_scratch_bytes(B, N))
    assert L.dfir_lam_backward(rev[N - 1].data_ptr(), -per_map, fsc.data_ptr(), float(gamma), gd.data_ptr(),
                               drev[N - 1].data_ptr(), -per_map, dgamma.data_ptr(), sc.data_ptr(), sc.numel(), N, B, H * W, 64,
                               G.stream()) == 0
    G.sync()
    got = torch.stack([G.to_nchw(drev[N - 1 - n]) for n in range(N)], dim=1)     # [B][N][64][H][W]
    # fp32 Gram sums over up to 65 536 elements feed a softmax: 2e-4 of the gradient norm (measured 7e-5 at 11 x 32x32)
    assert _relg(got, xr.grad) <= 2e-4
    assert abs(float(dgamma.cpu()) - float(gr.grad)) <= 2e-4 * max(1.0, abs(float(gr.grad)))


@pytest.mark.parametrize("shape", [(2, 6, 9), (1, 1, 1), (1, 17, 5)])
def test_channel_spatial_attention_backward(shape):
    B, H, W = shape
    g = torch.Generator().manual_seed(H * 3 + W + 7)
    x = torch.randn(B, 64, H, W, generator=g)
    sd = {"c.conv.weight": torch.randn(1, 1, 3, 3, 3, generator=g) * 0.3, "c.conv.bias": torch.randn(1, generator=g) * 0.1,
          "c.gamma": torch.tensor([0.6])}
    leaves = {k: v.double().requires_grad_(True) for k, v in sd.items()}
    xr = x.double().requires_grad_(True)
    gout = torch.randn(B, 64, H, W, generator=g)
    O.csam(xr, leaves, "c").backward(gout.double())
    L = G.lib()
    xd, gd = G.nhwc_f32(x), G.nhwc_f32(gout)
    w27 = sd["c.conv.weight"].reshape(-1).contiguous().cuda()
    nan = lambda *s: torch.full(s, float("nan"), device="cuda")
    dx, dw, db, dg = nan(B, H, W, 64), nan(27), nan(1), nan(1)
    sc = _scratch(L.dfir_csam_backward_scratch_bytes(B, H, W, 64))
    assert L.dfir_csam_backward(xd.data_ptr(), gd.data_ptr(), w27.data_ptr(), float(sd["c.conv.bias"]), float(sd["c.gamma"]),
                                dx.data_ptr(), dw.data_ptr(), db.data_ptr(), dg.data_ptr(), sc.data_ptr(), sc.numel(), B, H, W,
                                64, G.stream()) == 0
    G.sync()
    assert _relg(G.to_nchw(dx), xr.grad) <= 2e-5
    assert _relg(dw.cpu(), leaves["c.conv.weight"].grad.reshape(-1)) <= 2e-5
    assert _relg(db.cpu(), leaves["c.conv.bias"].grad) <= 2e-5
    assert _relg(dg.cpu(), leaves["c.gamma"].grad) <= 2e-5


@pytest.mark.parametrize("B,R", [(3, 4), (1, 8)])
def test_soca_mlp_forward_and_backward(B, R):
    g = torch.Generator().manual_seed(B * 10 + R)
    S = torch.randn(B, 64, 64, generator=g)
    w1, b1 = torch.randn(R, 64, generator=g) * 0.3, torch.randn(R, generator=g) * 0.1
    w2, b2 = torch.randn(64, R, generator=g) * 0.3, torch.randn(64, generator=g) * 0.1
    leaves = [t.double().requires_grad_(True) for t in (S, w1, b1, w2, b2)]
    v = leaves[0].mean(dim=1)
    want = torch.sigmoid(torch.relu(v @ leaves[1].t() + leaves[2]) @ leaves[3].t() + leaves[4])
    gs = torch.randn(B, 64, generator=g)
    want.backward(gs.double())
    L = G.lib()
    mlp = torch.cat([t.reshape(-1) for t in (w1, b1, w2, b2)]).contiguous().cuda()
    Sd, gsd = S.cuda(), gs.cuda()
    sv = torch.full((B, 64), float("nan"), device="cuda")
    assert L.dfir_soca_mlp(Sd.data_ptr(), mlp.data_ptr(), R, sv.data_ptr(), B, G.stream()) == 0
    dS = torch.full((B, 64, 64), float("nan"), device="cuda")
    dm = torch.full_like(mlp, float("nan"))
    sc = _scratch(L.dfir_soca_mlp_backward_scratch_bytes(B, R))
    assert L.dfir_soca_mlp_backward(Sd.data_ptr(), gsd.data_ptr(), mlp.data_ptr(), R, dS.data_ptr(), dm.data_ptr(), sc.data_ptr(),
                                    sc.numel(), B, G.stream()) == 0
    G.sync()
    assert _rel(sv.cpu(), want.detach()) <= 1e-5
    assert _relg(dS.cpu(), leaves[0].grad) <= 1e-5
    assert _relg(dm.cpu(), torch.cat([t.grad.reshape(-1) for t in leaves[1:]])) <= 1e-5


@pytest.mark.parametrize("B,HW,C", [(2, 37, 64), (1, 4096, 64), (3, 5, 128), (1, 9, 704)])
def test_channel_dot(B, HW, C):
    g = torch.Generator().manual_seed(HW + C)
    a, b = torch.randn(B, HW, C, generator=g), torch.randn(B, HW, C, generator=g)
    want = (a.double() * b.double()).sum(dim=1)
    L = G.lib()
    ad, bd = a.cuda(), b.cuda()
    out = torch.full((B, C), float("nan"), device="cuda")
    tot = torch.full((1,), 2.5, device="cuda")
    sc = _scratch(L.dfir_channel_dot_scratch_bytes(B, C))
    assert L.dfir_channel_dot(ad.data_ptr(), bd.data_ptr(), out.data_ptr(), tot.data_ptr(), 1, sc.data_ptr(), sc.numel(), B, HW,
                              C, G.stream()) == 0
    G.sync()
    assert _relg(out.cpu(), want) <= 1e-5
    assert abs(float(tot.cpu()) - 2.5 - float(want.sum())) <= 1e-4 * float(want.abs().sum())


def test_conv_data_gradient_through_the_transposed_fp32_pack():
    """dfir_pack_conv3x3_f32_ex(transpose=1) + dfir_conv3x3_f32 = data gradient of a 3x3 conv with Cin != Cout (the fusion
    convs of Q-HAN: 704 -> 64 and 128 -> 64)"""
    import torch.nn.functional as F
    B, H, W, cin, cout = 1, 6, 7, 128, 64
    g = torch.Generator().manual_seed(11)
    w = torch.randn(cout, cin, 3, 3, generator=g) * 0.1
    x = torch.randn(B, cin, H, W, generator=g).double().requires_grad_(True)
    gout = torch.randn(B, cout, H, W, generator=g)
    F.conv2d(x, w.double(), padding=1).backward(gout.double())
    L = G.lib()
    wT = torch.empty(9 * cin * cout, device="cuda")
    wd = w.contiguous().cuda()
    assert L.dfir_pack_conv3x3_f32_ex(wd.data_ptr(), wT.data_ptr(), cout, cin, 1, G.stream()) == 0
    gd = G.nhwc_f32(gout)
    dx = torch.full((B, H, W, cin), float("nan"), device="cuda")
    assert L.dfir_conv3x3_f32(gd.data_ptr(), wT.data_ptr(), None, None, dx.data_ptr(), B, H, W, cout, cin, 0, 1, 0,
                              G.stream()) == 0
    G.sync()
    assert _relg(G.to_nchw(dx), x.grad) <= 1e-5
