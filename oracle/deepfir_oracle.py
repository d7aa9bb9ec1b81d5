"""CPU restatement of the Deep-FIR hot path (TEST INFRASTRUCTURE ONLY — see oracle/__init__.py).

Every function restates one reference function as plain fp32 (or fp64) torch-CPU arithmetic driven
directly by a ``state_dict`` with the reference's key grammar, and cites the reference file:line it
follows (paths relative to ``/root/reference/Code/SISR/models``).  The network structure (number of
groups / blocks, which blocks own a meta-attention ``q_node``, the upsampler depth) is inferred from
the state_dict keys, so the same functions check any configuration the reference can build.

Parity pinning: the reference holds NO tests or golden vectors for this path, so the
restatement is pinned against outputs of the reference itself, generated in the build container by
``oracle/make_golden.py`` and committed under ``tests/golden/`` (see tests/test_oracle_golden.py).
``oracle/np_ops.py`` additionally pins the conv / pixel-shuffle / pooling primitives against an
independent numpy-float64 restatement.

``numerics`` (optional) lets tests emulate the storage-rounding policy of the bf16 GPU path on the
CPU (round conv operands to bf16, keep the residual stream fp32) to derive tolerances.
"""
from __future__ import annotations

import math
import re
from typing import Dict, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]


class Numerics:
    """Rounding policy hooks.  Default: exact fp32 (the reference's arithmetic)."""

    def __init__(self, conv_operand_dtype: Optional[torch.dtype] = None):
        self.conv_operand_dtype = conv_operand_dtype

    def q(self, t: Tensor) -> Tensor:
        """Round a conv operand (activation or weight) to the storage dtype and back."""
        if self.conv_operand_dtype is None:
            return t
        return t.to(self.conv_operand_dtype).to(t.dtype)


EXACT = Numerics()


# --------------------------------------------------------------------------------------------
# primitives
# --------------------------------------------------------------------------------------------
def conv3x3(x: Tensor, sd: SD, prefix: str, nm: Numerics = EXACT) -> Tensor:
    """`default_conv`: 3x3, stride 1, zero pad 1, bias (advanced/common.py:5-8)."""
    w = sd[prefix + ".weight"]
    b = sd.get(prefix + ".bias")
    return F.conv2d(nm.q(x), nm.q(w), b, padding=w.shape[-1] // 2)


def fc(v: Tensor, sd: SD, prefix: str) -> Tensor:
    """1x1 conv on a (B,C,1,1) vector == fully connected layer (q_layer.py:31, architectures.py:69-74)."""
    w = sd[prefix + ".weight"]
    return F.conv2d(v, w, sd.get(prefix + ".bias"))


def pixel_shuffle(x: Tensor, r: int) -> Tensor:
    """nn.PixelShuffle: out[b,c,h*r+i,w*r+j] = in[b, c*r*r + i*r + j, h, w] (advanced/common.py:30,36)."""
    b, c, h, w = x.shape
    x = x.reshape(b, c // (r * r), r, r, h, w).permute(0, 1, 4, 2, 5, 3)
    return x.reshape(b, c // (r * r), h * r, w * r)


def upsampler(x: Tensor, sd: SD, prefix: str, scale: int, nm: Numerics = EXACT) -> Tensor:
    """`Upsampler` (advanced/common.py:20-45): scale 2^n -> n x [conv C->4C, PixelShuffle(2)];
    scale 3 -> conv C->9C, PixelShuffle(3).  No activation (act=False in all Q-nets)."""
    if scale & (scale - 1) == 0:
        for i in range(int(math.log2(scale))):
            x = pixel_shuffle(conv3x3(x, sd, "%s.%d" % (prefix, 2 * i), nm), 2)
    elif scale == 3:
        x = pixel_shuffle(conv3x3(x, sd, prefix + ".0", nm), 3)
    else:
        raise NotImplementedError
    return x


def _has(sd: SD, key: str) -> bool:
    return key in sd


def _count(sd: SD, pattern: str) -> int:
    """number of distinct integer indices matching regex `pattern` (one capture group)."""
    rx = re.compile(pattern)
    idx = set()
    for k in sd:
        m = rx.match(k)
        if m:
            idx.add(int(m.group(1)))
    return (max(idx) + 1) if idx else 0


# --------------------------------------------------------------------------------------------
# attention vectors
# --------------------------------------------------------------------------------------------
def para_ca_vector(meta: Tensor, sd: SD, prefix: str) -> Tensor:
    """ParaCALayer.attribute_integrator (attention_manipulators/q_layer.py:21-37, 39-41).
    The Sequential holds Conv2d layers at ascending indices, a ReLU after each non-final conv iff
    `nonlinearity`, and a final Sigmoid.  Conv indices step by 2 when ReLUs are present, by 1 when
    not, so the presence of index '1' as a conv distinguishes the two."""
    idxs = sorted(int(m.group(1)) for k in sd
                  for m in [re.match(re.escape(prefix) + r"\.attribute_integrator\.(\d+)\.weight$", k)] if m)
    relu = not (len(idxs) > 1 and idxs[1] == idxs[0] + 1)
    y = meta
    for n, i in enumerate(idxs):
        y = fc(y, sd, "%s.attribute_integrator.%d" % (prefix, i))
        if relu and n != len(idxs) - 1:
            y = F.relu(y)
    return torch.sigmoid(y)


def qca_vector(x: Tensor, attributes: Tensor, sd: SD, prefix: str, style: str) -> Tensor:
    """QCALayer.forward up to (not including) the `x * y` (attention_manipulators/architectures.py:105-125)."""
    y = x.mean(dim=(2, 3), keepdim=True)  # AdaptiveAvgPool2d(1), :55,107
    du = prefix + ".conv_du"
    if style in ("standard", "modulate"):
        y = torch.sigmoid(fc(F.relu(fc(y, sd, du + ".0")), sd, du + ".2"))
        if style == "modulate":
            y = y * attributes  # :109
    elif style in ("max_concat", "softmax"):
        y = torch.cat((y, attributes), dim=1)  # :111,119
        y = torch.sigmoid(fc(F.relu(fc(y, sd, du + ".0")), sd, du + ".2"))
        if style == "softmax":
            y = torch.softmax(y, dim=1)  # :100-101,120 (softmax AFTER the sigmoid)
    elif style == "mini_concat":
        y = fc(y, sd, prefix + ".pre_concat")  # :113
        y = torch.cat((y, attributes), dim=1)
        y = torch.sigmoid(fc(F.relu(y), sd, du + ".1"))  # Sequential(ReLU, Conv, Sigmoid) :77-81
    elif style == "extended_attention":
        for i in range(3):  # :116-117
            y = F.relu(fc(torch.cat((y, attributes), dim=1), sd, "%s.feature_convs.%d.0" % (prefix, i)))
        y = torch.sigmoid(fc(y, sd, prefix + ".final_conv.0"))  # :118
    else:
        raise NotImplementedError
    return y


def pa_map(x: Tensor, sd: SD, prefix: str) -> Tensor:
    """PALayer (attention_manipulators/architectures.py:13-26): sigma(conv1x1(relu(conv1x1(x))))."""
    return torch.sigmoid(fc(F.relu(fc(x, sd, prefix + ".pa.0")), sd, prefix + ".pa.2"))


# --------------------------------------------------------------------------------------------
# Q-RCAN
# --------------------------------------------------------------------------------------------
def qrcab(x: Tensor, meta: Tensor, sd: SD, prefix: str, style: str, nm: Numerics = EXACT) -> Tensor:
    """QRCAB.forward (attention_manipulators/architectures.py:172-180).  `res_scale` is stored but
    never applied by the reference."""
    res = conv3x3(F.relu(conv3x3(x, sd, prefix + ".body.0", nm)), sd, prefix + ".body.2", nm)
    res = res * qca_vector(res, meta, sd, prefix + ".final_body", style)
    if _has(sd, prefix + ".pa_node.pa.0.weight"):
        res = res * pa_map(res, sd, prefix + ".pa_node")
    if _has(sd, prefix + ".q_node.attribute_integrator.0.weight"):
        res = res * para_ca_vector(meta, sd, prefix + ".q_node")
    return res + x


def qresidual_group(x: Tensor, meta: Tensor, sd: SD, prefix: str, style: str, nm: Numerics = EXACT) -> Tensor:
    """QResidualGroup.forward (attention_manipulators/architectures.py:229-233)."""
    nblocks = _count(sd, re.escape(prefix) + r"\.body\.(\d+)\.body\.0\.weight$")
    res = x
    for b in range(nblocks):
        res = qrcab(res, meta, sd, "%s.body.%d" % (prefix, b), style, nm)
    return conv3x3(res, sd, prefix + ".final_body", nm) + x


def _tail_scale(sd: SD, prefix: str = "tail") -> int:
    w0 = sd[prefix + ".0.0.weight"]
    ratio = w0.shape[0] // w0.shape[1]
    if ratio == 9:
        return 3
    n = _count(sd, re.escape(prefix) + r"\.0\.(\d+)\.weight$")  # conv indices 0,2,4..
    return 2 ** ((n + 1) // 2)


def qrcan_forward(x: Tensor, meta: Tensor, sd: SD, style: str = "standard", nm: Numerics = EXACT) -> Tensor:
    """QRCAN.forward (attention_manipulators/architectures.py:309-316)."""
    ngroups = _count(sd, r"body\.(\d+)\.final_body\.weight$")
    h = conv3x3(x, sd, "head.0", nm)
    res = h
    for g in range(ngroups):
        res = qresidual_group(res, meta, sd, "body.%d" % g, style, nm)
    res = conv3x3(res, sd, "final_body", nm) + h
    scale = _tail_scale(sd)
    return conv3x3(upsampler(res, sd, "tail.0", scale, nm), sd, "tail.1", nm)


# --------------------------------------------------------------------------------------------
# Q-EDSR
# --------------------------------------------------------------------------------------------
def qedsr_forward(x: Tensor, meta: Tensor, sd: SD, res_scale: float = 0.1, nm: Numerics = EXACT) -> Tensor:
    """QEDSR.forward + ParamResBlock.forward (attention_manipulators/architectures.py:346-356, 391-399):
    r = conv2(relu(conv1(x))) * res_scale ; r = r * ParaCALayer(meta) ; r += x."""
    nblocks = _count(sd, r"body\.(\d+)\.body\.0\.weight$")
    h = conv3x3(x, sd, "head", nm)
    res = h
    for b in range(nblocks):
        p = "body.%d" % b
        r = conv3x3(F.relu(conv3x3(res, sd, p + ".body.0", nm)), sd, p + ".body.2", nm)
        r = r * res_scale
        r = r * para_ca_vector(meta, sd, p + ".attention_layer")
        res = r + res
    res = conv3x3(res, sd, "final_body", nm) + h
    scale = _tail_scale(sd)
    return conv3x3(upsampler(res, sd, "tail.0", scale, nm), sd, "tail.1", nm)


# --------------------------------------------------------------------------------------------
# non-meta baselines RCAN / EDSR (advanced/architectures.py) — SURVEY.md §8f rank 3
# --------------------------------------------------------------------------------------------
def rcan_forward(x: Tensor, sd: SD, nm: Numerics = EXACT) -> Tensor:
    """RCAN.forward (advanced/architectures.py:159-164) with RCAB (:71-74: conv-ReLU-conv-CALayer, `res += x`),
    CALayer (:30-33: avg-pool -> FC-ReLU-FC-sigmoid -> `x * y`) and ResidualGroup (:107-110).  The trunk tail conv
    is the last entry of `body`."""
    ngroups = _count(sd, r"body\.(\d+)\.body\.0\.body\.0\.weight$")
    h = conv3x3(x, sd, "head.0", nm)
    res = h
    for g in range(ngroups):
        gp = "body.%d" % g
        nblocks = _count(sd, re.escape(gp) + r"\.body\.(\d+)\.body\.0\.weight$")
        r = res
        for b in range(nblocks):
            p = "%s.body.%d.body" % (gp, b)
            t = conv3x3(F.relu(conv3x3(r, sd, p + ".0", nm)), sd, p + ".2", nm)
            y = t.mean(dim=(2, 3), keepdim=True)
            y = torch.sigmoid(fc(F.relu(fc(y, sd, p + ".3.conv_du.0")), sd, p + ".3.conv_du.2"))
            r = t * y + r
        res = conv3x3(r, sd, "%s.body.%d" % (gp, nblocks), nm) + res
    res = conv3x3(res, sd, "body.%d" % ngroups, nm) + h
    scale = _tail_scale(sd)
    return conv3x3(upsampler(res, sd, "tail.0", scale, nm), sd, "tail.1", nm)


def edsr_forward(x: Tensor, sd: SD, res_scale: float = 0.1, nm: Numerics = EXACT) -> Tensor:
    """EDSR.forward (advanced/architectures.py:221-226) with ResBlock (advanced/common.py:68-72:
    `res = body(x).mul(res_scale); res += x`)."""
    nblocks = _count(sd, r"body\.(\d+)\.body\.0\.weight$")
    h = conv3x3(x, sd, "head.0", nm)
    res = h
    for b in range(nblocks):
        p = "body.%d.body" % b
        res = conv3x3(F.relu(conv3x3(res, sd, p + ".0", nm)), sd, p + ".2", nm) * res_scale + res
    res = conv3x3(res, sd, "body.%d" % nblocks, nm) + h
    scale = _tail_scale(sd)
    return conv3x3(upsampler(res, sd, "tail.0", scale, nm), sd, "tail.1", nm)


# --------------------------------------------------------------------------------------------
# Q-SAN pieces (SOCA second-order attention; non-local region attention)
# --------------------------------------------------------------------------------------------
def covpool(x: Tensor) -> Tensor:
    """Covpool.forward (advanced/mpncov.py:12-33): Sigma = X (I/M - 11^T/M^2) X^T, restated without
    materialising the MxM centring matrix: Sigma = X X^T / M - (X1)(X1)^T / M^2."""
    b, c, h, w = x.shape
    m = h * w
    xf = x.reshape(b, c, m)
    s = xf.sum(dim=2, keepdim=True)
    return xf.bmm(xf.transpose(1, 2)) / m - s.bmm(s.transpose(1, 2)) / (m * m)


def sqrtm_ns(a: Tensor, iters: int = 5) -> Tensor:
    """Sqrtm.forward, Newton-Schulz iteration (advanced/mpncov.py:49-76), iterN >= 2 branch."""
    b, d, _ = a.shape
    eye3 = 3.0 * torch.eye(d, dtype=a.dtype).expand(b, d, d)
    norm = a.diagonal(dim1=1, dim2=2).sum(dim=1)  # (1/3)*sum(x*I3) == trace, :57
    A = a / norm.view(b, 1, 1)
    ZY = 0.5 * (eye3 - A)
    Y = A.bmm(ZY)
    Z = ZY
    for _ in range(1, iters - 1):
        ZY = 0.5 * (eye3 - Z.bmm(Y))
        Y, Z = Y.bmm(ZY), ZY.bmm(Z)
    ZY = 0.5 * Y.bmm(eye3 - Z.bmm(Y))
    return ZY * torch.sqrt(norm).view(b, 1, 1)


def soca_vector(x: Tensor, sd: SD, prefix: str) -> Tensor:
    """SOCA.forward up to the final `y*x` (advanced/SAN_blocks.py:261-300): centre-crop to 1000 when a
    side is >= 1000, cov-pool, NS sqrt (5 it), mean over dim 1, FC-ReLU-FC-sigmoid."""
    _, _, h, w = x.shape
    h1 = w1 = 1000
    if h < h1 and w < w1:
        xs = x
    elif h < h1 and w > w1:
        W = (w - w1) // 2
        xs = x[:, :, :, W:(W + w1)]
    elif w < w1 and h > h1:
        H = (h - h1) // 2
        xs = x[:, :, H:H + h1, :]
    else:
        H = (h - h1) // 2
        W = (w - w1) // 2
        xs = x[:, :, H:(H + h1), W:(W + w1)]
    cov_sqrt = sqrtm_ns(covpool(xs), 5)
    v = cov_sqrt.mean(dim=1).reshape(x.shape[0], x.shape[1], 1, 1)
    return torch.sigmoid(fc(F.relu(fc(v, sd, prefix + ".conv_du.0")), sd, prefix + ".conv_du.2"))


def nonlocal_region(x: Tensor, sd: SD, prefix: str) -> Tensor:
    """_NonLocalBlockND._embedded_gaussian, 2-D (advanced/SAN_blocks.py:104-148) with the always-on
    MaxPool2d(2) after g and phi (`sub_sample` is shadowed at :36-40)."""
    b, c, h, w = x.shape
    g = F.max_pool2d(fc(x, sd, prefix + ".g.0"), 2)
    phi = F.max_pool2d(fc(x, sd, prefix + ".phi.0"), 2)
    theta = fc(x, sd, prefix + ".theta")
    ic = theta.shape[1]
    g_x = g.reshape(b, ic, -1).permute(0, 2, 1)
    theta_x = theta.reshape(b, ic, -1).permute(0, 2, 1)
    phi_x = phi.reshape(b, ic, -1)
    f = torch.softmax(theta_x.bmm(phi_x), dim=-1)
    y = f.bmm(g_x).permute(0, 2, 1).reshape(b, ic, h, w)
    return fc(y, sd, prefix + ".W") + x


def nonlocal_ca(x: Tensor, sd: SD, prefix: str) -> Tensor:
    """Nonlocal_CA.forward (advanced/SAN_blocks.py:314-336): 2x2 region split, shared weights."""
    _, _, H, W = x.shape
    H1, W1 = int(H / 2), int(W / 2)
    out = torch.zeros_like(x)
    p = prefix + ".non_local"
    out[:, :, :H1, :W1] = nonlocal_region(x[:, :, :H1, :W1], sd, p)
    out[:, :, H1:, :W1] = nonlocal_region(x[:, :, H1:, :W1], sd, p)
    out[:, :, :H1, W1:] = nonlocal_region(x[:, :, :H1, W1:], sd, p)
    out[:, :, H1:, W1:] = nonlocal_region(x[:, :, H1:, W1:], sd, p)
    return out


def qsan_forward(x: Tensor, meta: Tensor, sd: SD, nm: Numerics = EXACT) -> Tensor:
    """QSAN.forward (attention_manipulators/architectures.py:447-467) with QLSRAG / QRB
    (attention_manipulators/qsan_blocks.py:29-34, 66-84)."""
    ngroups = _count(sd, r"RG\.(\d+)\.conv_last\.weight$")
    h = conv3x3(x, sd, "head.0", nm)
    xx = nonlocal_ca(h, sd, "non_local")
    residual = xx
    for g in range(ngroups):
        p = "RG.%d" % g
        nblocks = _count(sd, re.escape(p) + r"\.rcab\.(\d+)\.conv_first\.0\.weight$")
        flow = xx
        for b in range(nblocks):
            q = "%s.rcab.%d" % (p, b)
            y = conv3x3(F.relu(conv3x3(flow, sd, q + ".conv_first.0", nm)), sd, q + ".conv_first.2", nm)
            if _has(sd, q + ".q_layer.attribute_integrator.0.weight"):  # absent in the non-meta SAN (RB, SAN_blocks.py:359-363)
                y = y * para_ca_vector(meta, sd, q + ".q_layer")
            flow = y + flow
        flow = flow * soca_vector(flow, sd, p + ".soca")
        flow = conv3x3(flow, sd, p + ".conv_last", nm)
        xx = (xx + flow) + sd["gamma"] * residual
    res = nonlocal_ca(xx, sd, "non_local") + h
    scale = _tail_scale(sd)
    return conv3x3(upsampler(res, sd, "tail.0", scale, nm), sd, "tail.1", nm)


# --------------------------------------------------------------------------------------------
# Q-HAN pieces
# --------------------------------------------------------------------------------------------
def lam(x5: Tensor, gamma: Tensor) -> Tensor:
    """LAM_Module.forward (advanced/HAN_blocks.py:24-37)."""
    b, n, c, h, w = x5.shape
    q = x5.reshape(b, n, -1)
    energy = q.bmm(q.transpose(1, 2))
    energy_new = energy.max(dim=-1, keepdim=True)[0] - energy
    att = torch.softmax(energy_new, dim=-1)
    out = att.bmm(q).reshape(b, n, c, h, w)
    return (gamma * out + x5).reshape(b, n * c, h, w)


def csam(x: Tensor, sd: SD, prefix: str) -> Tensor:
    """CSAM_Module.forward (advanced/HAN_blocks.py:59-76)."""
    out = torch.sigmoid(F.conv3d(x.unsqueeze(1), sd[prefix + ".conv.weight"], sd[prefix + ".conv.bias"], padding=1))
    out = (sd[prefix + ".gamma"] * out).reshape(x.shape)
    return x * out + x


def _rcan_group(res: Tensor, sd: SD, gp: str, nm: Numerics) -> Tensor:
    """ResidualGroup of the non-meta RCAN / HAN (advanced/architectures.py:94-110): RCABs, conv, `+= x`."""
    nblocks = _count(sd, re.escape(gp) + r"\.body\.(\d+)\.body\.0\.weight$")
    r = res
    for b in range(nblocks):
        p = "%s.body.%d.body" % (gp, b)
        t = conv3x3(F.relu(conv3x3(r, sd, p + ".0", nm)), sd, p + ".2", nm)
        y = t.mean(dim=(2, 3), keepdim=True)
        y = torch.sigmoid(fc(F.relu(fc(y, sd, p + ".3.conv_du.0")), sd, p + ".3.conv_du.2"))
        r = t * y + r
    return conv3x3(r, sd, "%s.body.%d" % (gp, nblocks), nm) + res


def san_forward(x: Tensor, sd: SD, nm: Numerics = EXACT) -> Tensor:
    """SAN.forward (advanced/architectures.py:288-311): Q-SAN's data flow without the meta-attention scale."""
    return qsan_forward(x, None, sd, nm)


def qhan_forward(x: Tensor, meta: Tensor, sd: SD, nm: Numerics = EXACT) -> Tensor:
    """QHAN.forward (attention_manipulators/architectures.py:514-540); with meta = None the non-meta HAN.forward
    (advanced/architectures.py:351-377), whose groups carry RCAN's key grammar."""
    meta_net = any(k.endswith(".final_body.weight") for k in sd)
    ngroups = _count(sd, r"body\.(\d+)\.final_body\.weight$") if meta_net else \
        _count(sd, r"body\.(\d+)\.body\.0\.body\.0\.weight$")
    h = conv3x3(x, sd, "head.0", nm)
    res = h
    stack = []
    for g in range(ngroups):
        res = qresidual_group(res, meta, sd, "body.%d" % g, "standard", nm) if meta_net else \
            _rcan_group(res, sd, "body.%d" % g, nm)
        stack.insert(0, res)
    res = conv3x3(res, sd, "body.%d" % ngroups, nm)
    stack.insert(0, res)
    out1 = res
    la = lam(torch.stack(stack, dim=1), sd["la.gamma"])
    out2 = F.conv2d(nm.q(la), nm.q(sd["last_conv.weight"]), sd["last_conv.bias"], padding=1)
    out1 = csam(out1, sd, "csa")
    res = F.conv2d(nm.q(torch.cat([out1, out2], 1)), nm.q(sd["last.weight"]), sd["last.bias"], padding=1)
    res = res + h
    scale = _tail_scale(sd)
    return conv3x3(upsampler(res, sd, "tail.0", scale, nm), sd, "tail.1", nm)


# --------------------------------------------------------------------------------------------
# host-side metadata formatting (QModel.generate_channels / QRCANHandler.scale_qpi)
# --------------------------------------------------------------------------------------------
def generate_channels(metadata: Tensor, keys, wanted, num_metadata: int) -> Tensor:
    """QModel.generate_channels (attention_manipulators/__init__.py:30-51), vectorised: select the
    metadata columns whose key[0] is in `wanted` (all columns when 'all' is wanted), -> (B,M,1,1) fp32."""
    if metadata is None:
        raise RuntimeError("Metadata needs to be specified for this network to run properly.")
    md = torch.as_tensor(metadata)
    if "all" in wanted:
        mask = [True] * num_metadata
    else:
        mask = [key[0] in wanted for key in keys]
    if len(keys) == 1:
        sel = md.reshape(md.shape[0], -1)
    else:
        sel = md[:, torch.tensor(mask)]
    out = torch.ones(md.shape[0], num_metadata) * sel.to(torch.float32)
    return out.unsqueeze(2).unsqueeze(3)


def scale_qpi(qpi: Tensor, n_feats: int = 64, min_mu: float = -0.2, max_mu: float = 0.8,
              clamp: bool = False, sig: float = 0.2) -> Tensor:
    """QRCANHandler.scale_qpi / gaussian (attention_manipulators/handlers.py:42-54)."""
    import numpy as np
    base = np.linspace(0, 1, n_feats)
    mu = (qpi * (max_mu - min_mu) + min_mu).reshape(qpi.shape[0], -1).double().numpy()
    g = (1 / (np.sqrt(2 * np.pi) * sig)) * np.exp(-np.power(base[None, :] - mu, 2.0) / (2 * sig ** 2))
    out = torch.from_numpy(g).to(torch.float32)
    if clamp:
        out = out.clamp(0, 1)
    return out.unsqueeze(2).unsqueeze(3)


def psnr(a: Tensor, b: Tensor, max_value: float = 1.0) -> float:
    """sr_tools/metrics.py:6-17 — 20*log10(max/sqrt(mse)); 100 for identical images."""
    mse = torch.mean((a.double() - b.double()) ** 2).item()
    if mse == 0:
        return 100.0
    return 20 * math.log10(max_value / math.sqrt(mse))
