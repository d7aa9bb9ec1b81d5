"""-m gpu: operator-level parity of the CUDA kernels (through the C ABI) against the CPU oracle."""
import ctypes as C

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import deepfir_oracle as O
from tests import gpu_util as G

pytestmark = pytest.mark.gpu


def _rand_case(B, H, W, cout=64, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(B, 64, H, W, generator=g) * 2 - 1
    w = (torch.rand(cout, 64, 3, 3, generator=g) * 2 - 1) / 24.0
    b = torch.rand(cout, generator=g) - 0.5
    return x, w, b


def _ref_conv(x, w, b):
    """fp32 conv of bf16-rounded operands == what bf16 tensor cores with fp32 accumulation compute."""
    return F.conv2d(G.bf16_round(x), G.bf16_round(w), b, padding=1)


# (B,H,W): full 128-wide rows, narrow image, two segments with a ragged tail, single row, many images
# ... and row bands that run across image boundaries with several rows per CTA (TMEM row-slot recycling)
SHAPES = [(2, 16, 128), (1, 9, 40), (1, 5, 200), (3, 1, 7), (5, 3, 128), (1, 130, 128), (3, 160, 128), (40, 9, 64),
          (2, 75, 300)]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("epi", [0, 1])
def test_conv_tc_bias_relu(shape, epi):
    B, H, W = shape
    x, w, b = _rand_case(B, H, W, seed=H * W)
    ref = _ref_conv(x, w, b)
    if epi == 1:
        ref = F.relu(ref)
    out, _, _ = G.conv_tc(G.nhwc_bf16(x), G.pack_bf16(w), b, epi)
    got = G.to_nchw(out)
    assert torch.isfinite(got).all()
    # output is stored as bf16: half an ulp = 2^-9 relative, plus fp32 accumulation-order noise
    assert torch.allclose(got, ref, rtol=2 ** -7, atol=2e-3), G.max_norm_err(got, ref)


def test_conv_tc_dx_shifted_views_one_hot_taps():
    """Hardware bring-up record (tools/diag_conv.py, profiles/r01_descriptor_bringup.md): the dx-shifted A
    views start 128/256 B into a swizzle atom; B200 takes the swizzle phase from the absolute smem address,
    so the descriptor base_offset stays 0.  One-hot taps make any chunk scrambling exact and visible."""
    H, W = 4, 128
    for dy, dx in ((1, 0), (1, 2), (0, 1), (2, 2)):
        w = torch.zeros(64, 64, 3, 3)
        w[torch.arange(64), torch.arange(64), dy, dx] = 1.0
        x = torch.arange(64).float().reshape(1, 64, 1, 1) + torch.arange(W).float().reshape(1, 1, 1, W) / 128 \
            + torch.zeros(1, 64, H, W)
        x = G.bf16_round(x)
        ref = F.conv2d(x, w, None, padding=1)
        out, _, _ = G.conv_tc(G.nhwc_bf16(x), G.pack_bf16(w), torch.zeros(64), 0)
        assert torch.equal(G.to_nchw(out), G.bf16_round(ref)), (dy, dx)


def test_conv_tc_pool_rows():
    B, H, W = 2, 7, 150
    x, w, b = _rand_case(B, H, W, seed=11)
    ref = _ref_conv(x, w, b)
    out, _, pool = G.conv_tc(G.nhwc_bf16(x), G.pack_bf16(w), b, 2)
    assert torch.allclose(G.to_nchw(out), ref, rtol=2 ** -7, atol=2e-3)
    # pool_rows[b][seg][y][c] = sum over the segment's valid pixels of the fp32 conv output
    rows = pool.cpu()  # [B][nseg][H][64]
    want0 = ref[:, :, :, :128].sum(dim=3).permute(0, 2, 1)
    want1 = ref[:, :, :, 128:].sum(dim=3).permute(0, 2, 1)
    assert torch.allclose(rows[:, 0], want0, rtol=1e-4, atol=1e-3)
    assert torch.allclose(rows[:, 1], want1, rtol=1e-4, atol=1e-3)


def test_conv_tc_skip_fp32_stream():
    B, H, W = 2, 5, 128
    x, w, b = _rand_case(B, H, W, seed=5)
    skip = torch.randn(B, 64, H, W)
    ref = _ref_conv(x, w, b) + skip
    out, o32, _ = G.conv_tc(G.nhwc_bf16(x), G.pack_bf16(w), b, 3, skip=G.nhwc_f32(skip), want_f32=True)
    assert G.max_norm_err(G.to_nchw(o32), ref) < 1e-5
    assert torch.allclose(G.to_nchw(out), ref, rtol=2 ** -7, atol=2e-3)


def test_conv_tc_tail_nchw():
    B, H, W = 1, 6, 256
    x, w, b = _rand_case(B, H, W, cout=3, seed=9)
    ref = _ref_conv(x, w, b)
    b16 = torch.zeros(16)
    b16[:3] = b
    _, o32, _ = G.conv_tc(G.nhwc_bf16(x), G.pack_bf16(w, nt_rows=16), b16, 4, cout=3)
    assert G.max_norm_err(o32.cpu(), ref) < 1e-5


@pytest.mark.parametrize("shape", [(2, 16, 128), (3, 5, 40), (1, 7, 200), (9, 2, 20)])
@pytest.mark.parametrize("epi,style", [(1, "standard"), (3, "standard"), (1, "max_concat"), (3, "softmax"),
                                       (1, "mini_concat"), (1, "extended_attention"), (3, "modulate"), (1, "none")])
def test_conv_tc_fused_scale_residual_input(shape, epi, style):
    """IN_FUSED: s = QCALayer(mean r, attributes) * meta scale; x' = r * s + x written back as the new fp32
    stream; out = conv(x') (QRCAB `res * y`, `res += x` fused into the next conv;
    architectures.py:105-127,172-180)."""
    from deepfir_b200.qrcan import ChannelAttentionParams
    B, H, W = shape
    M = 10
    A = 64 if style == "modulate" else M
    torch.manual_seed(W + epi)
    g = torch.Generator().manual_seed(W + epi)
    r = G.bf16_round(torch.randn(B, 64, H, W, generator=g))
    x = torch.randn(B, 64, H, W, generator=g)
    attr = torch.rand(B, A, generator=g)
    sq = torch.rand(B, 64, generator=g)
    if style != "none":
        ca = ChannelAttentionParams(64, style, 16, M)
        sd = {"p." + k: v for k, v in ca.state_dict().items()}
        blob = torch.cat([t.detach().reshape(-1) for t in ca.flat_params()]).cuda()
        sv = O.qca_vector(r, attr.reshape(B, A, 1, 1), sd, "p", style).reshape(B, 64) * sq
    else:
        blob = None
        sv = 0.1 * sq
    _, w, b = _rand_case(1, 1, 1, seed=W)
    xp = r * sv.reshape(B, 64, 1, 1) + x
    ref = _ref_conv(xp, w, b)
    skip = torch.randn(B, 64, H, W, generator=g) if epi == 3 else None
    ref = F.relu(ref) if epi == 1 else ref + skip
    nseg = (W + 127) // 128
    pool = torch.stack([r[:, :, :, 128 * s_:128 * (s_ + 1)].sum(dim=3).permute(0, 2, 1) for s_ in range(nseg)], 1)
    pool_d = pool.contiguous().cuda()  # [B][nseg][H][64]
    r_d, x_d, w_d, b_d, attr_d, sq_d = G.nhwc_bf16(r), G.nhwc_f32(x), G.pack_bf16(w), b.cuda(), attr.cuda(), sq.cuda()
    x_out = torch.full((B, H, W, 64), float("nan"), device="cuda")
    out = torch.full((B, H, W, 64), float("nan"), device="cuda", dtype=torch.bfloat16)
    o32 = torch.full((B, H, W, 64), float("nan"), device="cuda") if epi == 3 else None
    sk_d = G.nhwc_f32(skip) if skip is not None else None
    rc = G.lib().dfir_conv3x3_c64_fused(
        r_d.data_ptr(), x_d.data_ptr(), x_out.data_ptr(), pool_d.data_ptr(), STYLE_ID[style],
        blob.data_ptr() if blob is not None else None, 4, M, A, attr_d.data_ptr(), sq_d.data_ptr(), 0.1,
        w_d.data_ptr(), b_d.data_ptr(), B, H, W, epi, out.data_ptr(),
        sk_d.data_ptr() if sk_d is not None else None, o32.data_ptr() if o32 is not None else None, G.stream())
    assert rc == 0
    G.sync()
    assert G.max_norm_err(G.to_nchw(x_out), xp) < 2e-6          # the fp32 stream: one fma per element
    assert torch.allclose(G.to_nchw(out), ref, rtol=2 ** -7, atol=4e-3)
    if o32 is not None:  # rare 1-ulp bf16 flips of x' (fma on the GPU vs mul+add in the oracle) cost ~1e-5
        assert G.max_norm_err(G.to_nchw(o32), ref) < 1e-4


@pytest.mark.parametrize("shape", [(2, 16, 128), (3, 5, 40), (1, 7, 200), (9, 2, 20), (2, 1, 1), (1, 1, 9), (1, 6, 1)])
@pytest.mark.parametrize("style", ["standard", "max_concat", "extended_attention"])
def test_pool_by_linearity_trio(shape, style):
    """conv1+stats -> ca_from_stats -> conv2 with scale+skip epilogue == QRCAB.forward
    (attention_manipulators/architectures.py:172-180) including degenerate 1-pixel-wide/high images."""
    from deepfir_b200.qrcan import ChannelAttentionParams
    B, H, W = shape
    M = A = 10
    torch.manual_seed(H * 31 + W)
    g = torch.Generator().manual_seed(H * 31 + W)
    x = torch.randn(B, 64, H, W, generator=g)
    xin = G.bf16_round(x)
    w1 = (torch.rand(64, 64, 3, 3, generator=g) - 0.5) / 12
    w2 = (torch.rand(64, 64, 3, 3, generator=g) - 0.5) / 12
    b1 = torch.rand(64, generator=g) - 0.5
    b2 = torch.rand(64, generator=g) - 0.5
    attr = torch.rand(B, A, generator=g)
    sq = torch.rand(B, 64, generator=g)
    ca = ChannelAttentionParams(64, style, 16, M)
    sd = {"p." + k: v for k, v in ca.state_dict().items()}
    blob = torch.cat([t.detach().reshape(-1) for t in ca.flat_params()]).cuda()
    # oracle: the block exactly as the reference computes it (on bf16-rounded conv operands)
    t = G.bf16_round(F.relu(_ref_conv(xin, w1, b1)))
    r = _ref_conv(t, w2, b2)
    sv = O.qca_vector(r, attr.reshape(B, A, 1, 1), sd, "p", style).reshape(B, 64) * sq
    want = r * sv.reshape(B, 64, 1, 1) + x
    nseg = (W + 127) // 128
    x_bf, x32 = G.nhwc_bf16(xin), G.nhwc_f32(x)
    w1p, w2p, b1d, b2d, attr_d, sq_d = G.pack_bf16(w1), G.pack_bf16(w2), b1.cuda(), b2.cuda(), attr.cuda(), sq.cuda()
    t_d = torch.empty(B, H, W, 64, device="cuda", dtype=torch.bfloat16)
    pool = torch.full((B, nseg, H, 64), float("nan"), device="cuda")
    cf = torch.full((B, H, 64), float("nan"), device="cuda")
    cl = torch.full((B, H, 64), float("nan"), device="cuda")
    svec = torch.full((B, 64), float("nan"), device="cuda")
    out32 = torch.full((B, H, W, 64), float("nan"), device="cuda")
    outbf = torch.empty(B, H, W, 64, device="cuda", dtype=torch.bfloat16)
    L = G.lib()
    assert L.dfir_conv3x3_c64_stats(x_bf.data_ptr(), w1p.data_ptr(), b1d.data_ptr(), B, H, W, t_d.data_ptr(),
                                    pool.data_ptr(), cf.data_ptr(), cl.data_ptr(), G.stream()) == 0
    assert L.dfir_ca_from_stats(pool.data_ptr(), cf.data_ptr(), cl.data_ptr(), w2p.data_ptr(), b2d.data_ptr(),
                                STYLE_ID[style], blob.data_ptr(), 4, M, A, attr_d.data_ptr(), sq_d.data_ptr(),
                                svec.data_ptr(), B, H, W, G.stream()) == 0
    assert L.dfir_conv3x3_c64_scale_skip(t_d.data_ptr(), w2p.data_ptr(), b2d.data_ptr(), B, H, W, svec.data_ptr(),
                                         x32.data_ptr(), out32.data_ptr(), outbf.data_ptr(), None, None, None, 0,
                                         None, 4, M, A, None, None, G.stream()) == 0
    # same, with the attention vector evaluated inside the kernel from the statistics
    out32b = torch.full((B, H, W, 64), float("nan"), device="cuda")
    outbfb = torch.empty(B, H, W, 64, device="cuda", dtype=torch.bfloat16)
    assert L.dfir_conv3x3_c64_scale_skip(t_d.data_ptr(), w2p.data_ptr(), b2d.data_ptr(), B, H, W, None,
                                         x32.data_ptr(), out32b.data_ptr(), outbfb.data_ptr(), pool.data_ptr(),
                                         cf.data_ptr(), cl.data_ptr(), STYLE_ID[style], blob.data_ptr(), 4, M, A,
                                         attr_d.data_ptr(), sq_d.data_ptr(), G.stream()) == 0
    G.sync()
    assert G.max_norm_err(out32b, out32) < 2e-6
    # conv1 output: equal up to rare 1-ulp bf16 flips from the fp32 summation order
    assert torch.allclose(G.to_nchw(t_d), t, rtol=2 ** -7, atol=1e-3)
    assert torch.allclose(svec.cpu(), sv, rtol=1e-3, atol=1e-5), (svec.cpu() - sv).abs().max()
    assert G.max_norm_err(G.to_nchw(out32), want) < 1e-3
    assert torch.equal(outbf.cpu(), out32.cpu().to(torch.bfloat16))
    # the same block on the bf16 hi + lo residual stream (x = hi + lo): result planes against the fp32-stream epilogue
    x_hi = x32.to(torch.bfloat16)
    x_lo = (x32 - x_hi.float()).to(torch.bfloat16)
    o_hi = torch.full((B, H, W, 64), float("nan"), device="cuda", dtype=torch.bfloat16)
    o_lo = torch.full((B, H, W, 64), float("nan"), device="cuda", dtype=torch.bfloat16)
    assert L.dfir_conv3x3_c64_scale_skip_hl(t_d.data_ptr(), w2p.data_ptr(), b2d.data_ptr(), B, H, W, None, x_hi.data_ptr(),
                                            x_lo.data_ptr(), o_hi.data_ptr(), o_lo.data_ptr(), pool.data_ptr(),
                                            cf.data_ptr(), cl.data_ptr(), STYLE_ID[style], blob.data_ptr(), 4, M, A,
                                            attr_d.data_ptr(), sq_d.data_ptr(), 0, G.stream()) == 0
    G.sync()
    got = o_hi.float() + o_lo.float()
    scale = out32.abs().max().item()
    assert (got - out32).abs().max().item() <= 2.0 ** -15 * scale, (got - out32).abs().max().item() / scale
    assert (o_hi.float() - out32).abs().max().item() <= 2.0 ** -8 * scale   # hi alone is the bf16 rounding of the stream
    # descending traversal (images and rows from the last to the first), in-kernel attention: same values
    d_hi, d_lo = torch.empty_like(o_hi), torch.empty_like(o_lo)
    assert L.dfir_conv3x3_c64_scale_skip_hl(t_d.data_ptr(), w2p.data_ptr(), b2d.data_ptr(), B, H, W, None, x_hi.data_ptr(),
                                            x_lo.data_ptr(), d_hi.data_ptr(), d_lo.data_ptr(), pool.data_ptr(),
                                            cf.data_ptr(), cl.data_ptr(), STYLE_ID[style], blob.data_ptr(), 4, M, A,
                                            attr_d.data_ptr(), sq_d.data_ptr(), 1, G.stream()) == 0
    G.sync()
    assert ((d_hi.float() + d_lo.float()) - out32).abs().max().item() <= 2.0 ** -15 * scale
    # in place, scale vector from memory, descending
    assert L.dfir_conv3x3_c64_scale_skip_hl(t_d.data_ptr(), w2p.data_ptr(), b2d.data_ptr(), B, H, W, svec.data_ptr(),
                                            x_hi.data_ptr(), x_lo.data_ptr(), x_hi.data_ptr(), x_lo.data_ptr(), None, None,
                                            None, 0, None, 4, M, A, None, None, 1, G.stream()) == 0
    G.sync()
    assert ((x_hi.float() + x_lo.float()) - out32).abs().max().item() <= 2.0 ** -15 * scale
    # ---- the same on the 8-BIT lo plane (the format of dfir_qrcan_forward): the stream value is a 24-bit float X,
    # bits(X) = (hi << 16) + (q << 8), hi = nearest bf16 (ties away from zero), q = int8
    # the lo plane stores a pixel's 64 bytes in accumulator-fragment order: byte cq * 16 + 2 n + e holds channel 8 n + 2 cq + e
    kk = torch.arange(64)
    chan_of_byte = (8 * ((kk % 16) // 2) + 2 * (kk // 16) + (kk % 2)).cuda()

    def enc8(v32):
        bits = v32.contiguous().view(torch.int32).to(torch.int64) & 0xFFFFFFFF
        t = bits + 0x80
        hb = (t + 0x8000) & 0xFFFF0000
        q = ((t - hb) >> 8) & 0xFF
        hi = (hb >> 16).to(torch.int32).to(torch.int16).view(torch.bfloat16)      # (wraps to the signed 16-bit pattern)
        return hi.contiguous(), q.to(torch.uint8).view(torch.int8)[..., chan_of_byte].contiguous()

    def dec8(hi, q):
        qc = torch.empty_like(q)
        qc[..., chan_of_byte] = q                                                   # back to channel order
        hb = (hi.contiguous().view(torch.int16).to(torch.int64) & 0xFFFF) << 16
        bits = (hb + (qc.to(torch.int64) << 8)) & 0xFFFFFFFF
        bits = torch.where(bits >= 2 ** 31, bits - 2 ** 32, bits).to(torch.int32)
        return bits.view(torch.float32)

    y_hi, y_q = enc8(x32)
    assert (dec8(y_hi, y_q) - x32).abs().max().item() <= 2.0 ** -16 * x32.abs().max().item()
    # the library's own codec operators agree with that definition, bit for bit
    c_hi = torch.empty_like(y_hi); c_q = torch.empty_like(y_q); c_x = torch.empty_like(x32)
    assert L.dfir_stream_encode_hl8(x32.data_ptr(), c_hi.data_ptr(), c_q.data_ptr(), x32.numel(), G.stream()) == 0
    assert L.dfir_stream_decode_hl8(c_hi.data_ptr(), c_q.data_ptr(), c_x.data_ptr(), x32.numel(), G.stream()) == 0
    G.sync()
    assert torch.equal(c_hi.view(torch.int16), y_hi.view(torch.int16)) and torch.equal(c_q, y_q)
    assert torch.equal(c_x, dec8(y_hi, y_q))
    for desc in (0, 1):
        p_hi = torch.full((B, H, W, 64), float("nan"), device="cuda", dtype=torch.bfloat16)
        p_q = torch.full((B, H, W, 64), 77, device="cuda", dtype=torch.int8)
        assert L.dfir_conv3x3_c64_scale_skip_hl8(t_d.data_ptr(), w2p.data_ptr(), b2d.data_ptr(), B, H, W, None, y_hi.data_ptr(),
                                                 y_q.data_ptr(), p_hi.data_ptr(), p_q.data_ptr(), pool.data_ptr(),
                                                 cf.data_ptr(), cl.data_ptr(), STYLE_ID[style], blob.data_ptr(), 4, M, A,
                                                 attr_d.data_ptr(), sq_d.data_ptr(), desc, G.stream()) == 0
        G.sync()
        assert (dec8(p_hi, p_q) - out32).abs().max().item() <= 2.0 ** -14 * scale, desc
        assert (p_hi.float() - out32).abs().max().item() <= 2.0 ** -8 * scale
    # the statistics as the network schedule passes them: the 64-bit fixed-point image sums of dfir_conv3x3_c64_stats_fx through
    # `pool_rows` with col_first == col_last == NULL (the per-row arrays and the fixed-point sums differ by the rounding of the
    # sums only: same result within the stream's precision)
    t_fx = torch.empty_like(t_d)
    ist = torch.zeros(B, 9, 64, device="cuda", dtype=torch.int64)
    assert L.dfir_conv3x3_c64_stats_fx(x_bf.data_ptr(), w1p.data_ptr(), b1d.data_ptr(), B, H, W, t_fx.data_ptr(), ist.data_ptr(),
                                       0, G.stream()) == 0
    G.sync()
    assert torch.equal(t_fx, t_d)
    for desc in (0, 1):
        p_hi = torch.full((B, H, W, 64), float("nan"), device="cuda", dtype=torch.bfloat16)
        p_q = torch.full((B, H, W, 64), 77, device="cuda", dtype=torch.int8)
        assert L.dfir_conv3x3_c64_scale_skip_hl8(t_d.data_ptr(), w2p.data_ptr(), b2d.data_ptr(), B, H, W, None, y_hi.data_ptr(),
                                                 y_q.data_ptr(), p_hi.data_ptr(), p_q.data_ptr(), ist.data_ptr(), None, None,
                                                 STYLE_ID[style], blob.data_ptr(), 4, M, A, attr_d.data_ptr(), sq_d.data_ptr(),
                                                 desc, G.stream()) == 0
        G.sync()
        assert (dec8(p_hi, p_q) - out32).abs().max().item() <= 2.0 ** -14 * scale + 1e-5 * scale, desc
    # in place, scale vector from memory
    assert L.dfir_conv3x3_c64_scale_skip_hl8(t_d.data_ptr(), w2p.data_ptr(), b2d.data_ptr(), B, H, W, svec.data_ptr(),
                                             y_hi.data_ptr(), y_q.data_ptr(), y_hi.data_ptr(), y_q.data_ptr(), None, None,
                                             None, 0, None, 4, M, A, None, None, 1, G.stream()) == 0
    G.sync()
    assert (dec8(y_hi, y_q) - out32).abs().max().item() <= 2.0 ** -14 * scale
    # group-conv use: no scale vector, in-place stream update
    assert L.dfir_conv3x3_c64_scale_skip(t_d.data_ptr(), w2p.data_ptr(), b2d.data_ptr(), B, H, W, None,
                                         x32.data_ptr(), x32.data_ptr(), outbf.data_ptr(), None, None, None, 0,
                                         None, 4, M, A, None, None, G.stream()) == 0
    G.sync()
    assert G.max_norm_err(G.to_nchw(x32), r + x) < 1e-3


@pytest.mark.parametrize("r", [2, 3])
def test_conv_tc_pixel_shuffle_fold(r):
    """conv C -> r^2 C + PixelShuffle(r) (advanced/common.py:20-45) as r^2 strided-store launches."""
    B, H, W = 1, 6, 40
    g = torch.Generator().manual_seed(1)
    x = torch.rand(B, 64, H, W, generator=g) - 0.5
    w = (torch.rand(64 * r * r, 64, 3, 3, generator=g) - 0.5) / 12
    b = torch.rand(64 * r * r, generator=g) - 0.5
    ref = O.pixel_shuffle(_ref_conv(x, w, b), r)
    out = torch.full((B, H * r, W * r, 64), float("nan"), dtype=torch.bfloat16, device="cuda")
    xin = G.nhwc_bf16(x)
    oW = W * r
    for s in range(r * r):
        i, j = divmod(s, r)
        bias_s = b.reshape(64, r * r)[:, s]
        view = out.reshape(-1)[(i * oW + j) * 64:]
        G.conv_tc(xin, G.pack_bf16(w, co_begin=s, co_stride=r * r), bias_s, 0, out=view,
                  strides=(r * 128, r * oW * 128, H * r * oW * 128))
    assert torch.allclose(G.to_nchw(out), ref, rtol=2 ** -7, atol=2e-3)


@pytest.mark.parametrize("cfg", [(2, 9, 13, 64, 64, 1, 1, 0), (1, 6, 10, 64, 256, 0, 2, 0), (1, 5, 9, 64, 3, 0, 1, 1),
                                 (1, 4, 6, 256, 256, 1, 1, 0), (1, 4, 5, 64, 576, 0, 3, 0)])
def test_conv_f32_simt(cfg):
    B, H, W, Cin, Cout, relu, ps, nchw = cfg
    g = torch.Generator().manual_seed(Cout + W)
    x = torch.rand(B, Cin, H, W, generator=g) - 0.5
    w = (torch.rand(Cout, Cin, 3, 3, generator=g) - 0.5) / 8
    b = torch.rand(Cout, generator=g) - 0.5
    skip = torch.randn(B, Cout, H, W) if (ps == 1 and not nchw) else None
    ref = F.conv2d(x, w, b, padding=1)
    if skip is not None:
        ref = ref + skip
    if relu:
        ref = F.relu(ref)
    if ps > 1:
        ref = O.pixel_shuffle(ref, ps)
    if nchw:
        out = torch.empty(B, Cout, H, W, device="cuda")
    else:
        out = torch.empty(B, H * ps, W * ps, Cout // (ps * ps), device="cuda")
    sk = G.nhwc_f32(skip) if skip is not None else None
    xd, wd, bd = G.nhwc_f32(x), G.pack_f32(w), b.cuda()
    rc = G.lib().dfir_conv3x3_f32(xd.data_ptr(), wd.data_ptr(), bd.data_ptr(),
                                  sk.data_ptr() if sk is not None else None, out.data_ptr(), B, H, W, Cin, Cout, relu,
                                  ps, nchw, G.stream())
    assert rc == 0
    G.sync()
    got = out.cpu() if nchw else G.to_nchw(out)
    assert G.max_norm_err(got, ref) < 1e-5


def test_head_conv():
    B, H, W = 2, 11, 17
    g = torch.Generator().manual_seed(2)
    x = torch.rand(B, 3, H, W, generator=g)
    w = (torch.rand(64, 3, 3, 3, generator=g) - 0.5) / 3
    b = torch.rand(64, generator=g) - 0.5
    ref = F.conv2d(x, w, b, padding=1)
    o32 = torch.empty(B, H, W, 64, device="cuda")
    obf = torch.empty(B, H, W, 64, device="cuda", dtype=torch.bfloat16)
    xd, wd, bd = x.cuda(), G.pack_f32(w), b.cuda()
    rc = G.lib().dfir_head_conv(xd.data_ptr(), wd.data_ptr(), bd.data_ptr(), o32.data_ptr(),
                                obf.data_ptr(), B, 3, H, W, 64, G.stream())
    assert rc == 0
    G.sync()
    assert G.max_norm_err(G.to_nchw(o32), ref) < 1e-6
    assert torch.equal(obf.cpu(), o32.cpu().to(torch.bfloat16))


@pytest.mark.parametrize("M,hid,C,relu", [(10, 32, 64, 1), (10, 128, 256, 0), (1, 32, 64, 1), (40, 52, 64, 1)])
def test_meta_attention(M, hid, C, relu):
    nblk, B = 5, 3
    g = torch.Generator().manual_seed(M)
    meta = torch.rand(B, M, generator=g)
    w1 = torch.randn(nblk, hid, M, generator=g) / 3
    b1 = torch.randn(nblk, hid, generator=g) / 3
    w2 = torch.randn(nblk, C, hid, generator=g) / 5
    b2 = torch.randn(nblk, C, generator=g) / 3
    en = torch.tensor([1, 0, 1, 1, 0], dtype=torch.int32)
    out = torch.empty(nblk, B, C, device="cuda")
    cu = [t.cuda() for t in (meta, w1, b1, w2, b2, en)]
    rc = G.lib().dfir_meta_attention(*[t.data_ptr() for t in cu[:5]], out.data_ptr(), nblk, B, M, hid, C, relu,
                                     cu[5].data_ptr(), 0.5, G.stream())
    assert rc == 0
    G.sync()
    for k in range(nblk):
        h = meta @ w1[k].t() + b1[k]
        if relu:
            h = F.relu(h)
        want = 0.5 * (torch.sigmoid(h @ w2[k].t() + b2[k]) if en[k] else torch.ones(B, C))
        assert torch.allclose(out[k].cpu(), want, rtol=1e-5, atol=1e-6)


STYLE_ID = {"none": 0, "standard": 1, "modulate": 2, "max_concat": 3, "softmax": 4, "mini_concat": 5,
            "extended_attention": 6}


@pytest.mark.parametrize("style", ["standard", "modulate", "max_concat", "softmax", "mini_concat",
                                   "extended_attention", "none"])
@pytest.mark.parametrize("r_bf16", [0, 1])
def test_ca_scale_residual_all_styles(style, r_bf16):
    """QCALayer (all styles) * meta scale + residual vs the oracle restatement."""
    from deepfir_b200.qrcan import ChannelAttentionParams
    B, H, W, Cc, M = 2, 6, 10, 64, 10
    torch.manual_seed(7)
    A = 64 if style == "modulate" else M
    r = torch.randn(B, Cc, H, W)
    if r_bf16:
        r = G.bf16_round(r)
    x = torch.randn(B, Cc, H, W)
    attr = torch.rand(B, A)
    sq = torch.rand(B, Cc)
    res_scale = 0.1 if style == "none" else 1.0
    if style != "none":
        ca = ChannelAttentionParams(Cc, style, 16, M)
        sd = {"p." + k: v for k, v in ca.state_dict().items()}
        blob = torch.cat([t.detach().reshape(-1) for t in ca.flat_params()]).cuda()
        y = O.qca_vector(r, attr.reshape(B, A, 1, 1), sd, "p", style)
        want = r * y * sq.reshape(B, Cc, 1, 1) + x
    else:
        blob = None
        want = r * res_scale * sq.reshape(B, Cc, 1, 1) + x
    r_dev = G.nhwc_bf16(r) if r_bf16 else G.nhwc_f32(r)
    pool = G.nhwc_f32(r).sum(dim=2).contiguous()  # [B][H][C] row sums
    x_dev = G.nhwc_f32(x)
    out = torch.empty_like(x_dev)
    obf = torch.empty(B, H, W, Cc, device="cuda", dtype=torch.bfloat16)
    attr_d, sq_d = attr.cuda(), sq.cuda()
    rc = G.lib().dfir_ca_scale_residual(r_dev.data_ptr(), r_bf16, x_dev.data_ptr(), pool.data_ptr(), H,
                                        STYLE_ID[style], blob.data_ptr() if blob is not None else None, Cc, 4, M, A,
                                        attr_d.data_ptr(), sq_d.data_ptr(), res_scale, out.data_ptr(),
                                        obf.data_ptr(), B, H, W, G.stream())
    assert rc == 0
    G.sync()
    assert G.max_norm_err(G.to_nchw(out), want) < 2e-6
    assert torch.equal(obf.cpu(), out.cpu().to(torch.bfloat16))


def test_pool_rows_f32():
    x = torch.randn(2, 64, 5, 9)
    out = torch.empty(2, 5, 64, device="cuda")
    xd = G.nhwc_f32(x)
    assert G.lib().dfir_pool_rows_f32(xd.data_ptr(), out.data_ptr(), 2, 5, 9, 64, G.stream()) == 0
    G.sync()
    assert torch.allclose(out.cpu(), x.sum(dim=3).permute(0, 2, 1), rtol=1e-5, atol=1e-5)


def test_bad_arguments_return_error_codes():
    L = G.lib()
    assert L.dfir_conv3x3_f32(None, None, None, None, None, 1, 4, 4, 3, 8, 0, 1, 0, None) == -1  # Cin % 4
    x = torch.zeros(1, 4, 4, 64, device="cuda", dtype=torch.bfloat16)
    assert L.dfir_conv3x3_c64(x.data_ptr(), 64, 0, x.data_ptr(), x.data_ptr(), 1, 4, 4, 2, 64, x.data_ptr(), 128, 512,
                              2048, None, None, None, 0, None) == -1  # pool epilogue without pool buffer
    assert L.dfir_conv3x3_c64(x.data_ptr(), 64, 0, x.data_ptr(), x.data_ptr(), 1, 4, 4, 0, 64, x.data_ptr(), 128, 512,
                              2048, None, None, None, 1, None) == -1  # reserved desc_mode
    assert L.dfir_conv3x3_c64_fused(x.data_ptr(), None, None, None, 0, None, 4, 10, 10, None, None, 1.0, x.data_ptr(),
                                    x.data_ptr(), 1, 4, 4, 1, x.data_ptr(), None, None, None) == -1  # no x_in


@pytest.mark.parametrize("shape", [(2, 16, 128), (3, 5, 40), (1, 7, 200), (2, 1, 64), (1, 128, 128), (5, 24, 96)])
@pytest.mark.parametrize("warp_autonomous", [0, 1])
def test_conv1_fixed_point_image_statistics(shape, warp_autonomous):
    """dfir_conv3x3_c64_stats_fx: t = relu(conv(x) + b) in bf16 plus the nine image sums of pool-by-linearity (total, first /
    last column, first / last row, four corners) as 64-bit fixed-point integers; both epilogues (group-synchronous and
    warp-autonomous); the sums of an image do not depend on the rest of the batch (integer accumulation)"""
    B, H, W = shape
    g = torch.Generator().manual_seed(H * 17 + W)
    x = G.bf16_round(torch.randn(B, 64, H, W, generator=g))
    w = (torch.rand(64, 64, 3, 3, generator=g) - 0.5) / 12
    b = torch.rand(64, generator=g) - 0.5
    t = F.relu(_ref_conv(x, w, b)).double()                       # [B][64][H][W]
    want = torch.stack([t.sum(dim=(2, 3)), t[:, :, :, 0].sum(dim=2), t[:, :, :, -1].sum(dim=2), t[:, :, 0, :].sum(dim=2),
                        t[:, :, -1, :].sum(dim=2), t[:, :, 0, 0], t[:, :, 0, -1], t[:, :, -1, 0], t[:, :, -1, -1]], dim=1)
    L = G.lib()
    xd, wp, bd = G.nhwc_bf16(x), G.pack_bf16(w), b.cuda()

    def run(xdev, nb):
        out = torch.empty(nb, H, W, 64, device="cuda", dtype=torch.bfloat16)
        ist = torch.zeros(nb, 9, 64, device="cuda", dtype=torch.int64)
        rc = L.dfir_conv3x3_c64_stats_fx(xdev.data_ptr(), wp.data_ptr(), bd.data_ptr(), nb, H, W, out.data_ptr(), ist.data_ptr(),
                                         warp_autonomous, G.stream())
        return rc, out, ist

    rc, out, ist = run(xd, B)
    if warp_autonomous and (W - 1) % 128 < 32 and ((W - 1) % 128) % 8 == 0:
        assert rc != 0        # one column accumulator per thread: such widths take the group-synchronous epilogue
        return
    assert rc == 0
    G.sync()
    assert torch.allclose(G.to_nchw(out), t.float(), rtol=2 ** -7, atol=1e-3)
    got = ist.cpu().double() / 2 ** 24
    # quantisation: 2^-13 per (thread, row) partial in the warp-autonomous form, 2^-25 per row otherwise; fp32 sums of a row
    tol = 1e-4 * want.abs().max().item() + H * W * 2.0 ** -13 / 4
    assert (got - want).abs().max().item() <= tol, (got - want).abs().max().item()
    # an image alone == the same image inside the batch, bit for bit
    rc1, _, ist1 = run(xd[B - 1:].contiguous(), 1)
    assert rc1 == 0
    G.sync()
    assert torch.equal(ist1[0].cpu(), ist[B - 1].cpu())
