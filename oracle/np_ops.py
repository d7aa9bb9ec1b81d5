"""Independent numpy-float64 restatement of the primitive ops (TEST INFRASTRUCTURE ONLY) used to pin
the torch-based oracle at small sizes: it shares no code with ATen."""
import numpy as np


def conv3x3_np(x, w, b=None):
    """3x3 cross-correlation, stride 1, zero pad 1 (advanced/common.py:5-8 -> nn.Conv2d semantics):
    out[n,o,y,x] = b[o] + sum_{i,dy,dx} w[o,i,dy,dx] * xpad[n,i,y+dy,x+dx]."""
    x = np.asarray(x, dtype=np.float64)
    w = np.asarray(w, dtype=np.float64)
    n, c, h, wd = x.shape
    xp = np.zeros((n, c, h + 2, wd + 2), dtype=np.float64)
    xp[:, :, 1:-1, 1:-1] = x
    out = np.zeros((n, w.shape[0], h, wd), dtype=np.float64)
    for dy in range(3):
        for dx in range(3):
            out += np.einsum("oi,nihw->nohw", w[:, :, dy, dx], xp[:, :, dy:dy + h, dx:dx + wd])
    if b is not None:
        out += np.asarray(b, dtype=np.float64).reshape(1, -1, 1, 1)
    return out


def pixel_shuffle_np(x, r):
    """out[b,c,h*r+i,w*r+j] = in[b,c*r*r+i*r+j,h,w] (nn.PixelShuffle; advanced/common.py:30)."""
    x = np.asarray(x)
    b, c, h, w = x.shape
    oc = c // (r * r)
    out = np.zeros((b, oc, h * r, w * r), dtype=x.dtype)
    for ch in range(oc):
        for i in range(r):
            for j in range(r):
                out[:, ch, i::r, j::r] = x[:, ch * r * r + i * r + j]
    return out


def avgpool_np(x):
    return np.asarray(x, dtype=np.float64).mean(axis=(2, 3), keepdims=True)


def covpool_np(x):
    """Literal restatement of Covpool.forward (advanced/mpncov.py:12-33) WITH the MxM centring matrix."""
    x = np.asarray(x, dtype=np.float64)
    b, c, h, w = x.shape
    m = h * w
    xf = x.reshape(b, c, m)
    ihat = np.full((m, m), -1.0 / m / m)
    ihat[np.diag_indices(m)] += 1.0 / m
    return np.einsum("bcm,mn,bdn->bcd", xf, ihat, xf)


def batch_blur_np(x, kernels):
    """Literal restatement of BatchBlur.forward (Code/sr_tools/gaussian_utils.py:346-368): reflection pad (l//2 on the low
    side, l//2 or l//2 - 1 on the high side), then every plane of image b cross-correlated with kernels[b] (or the single
    shared kernel)."""
    x = np.asarray(x, dtype=np.float64)
    k = np.asarray(kernels, dtype=np.float64)
    b, c, h, w = x.shape
    l = k.shape[-1]
    lo, hi = l // 2, (l // 2 if l % 2 == 1 else l // 2 - 1)
    xp = np.pad(x, ((0, 0), (0, 0), (lo, hi), (lo, hi)), mode="reflect")
    out = np.zeros_like(x)
    for i in range(b):
        kk = k[i] if k.ndim == 3 else k
        for dy in range(l):
            for dx in range(l):
                out[i] += kk[dy, dx] * xp[i, :, dy:dy + h, dx:dx + w]
    return out


def pca_encode_np(kernels, pca):
    """PCAEncoder.__call__ (:333-343): [B, l*l] @ [l*l, k]"""
    k = np.asarray(kernels, dtype=np.float64)
    return k.reshape(k.shape[0], -1) @ np.asarray(pca, dtype=np.float64)
