"""Deterministic synthetic weights / inputs shared by the golden generator, the tests and the smoke
check (TEST INFRASTRUCTURE ONLY).  Uses numpy's frozen legacy ``RandomState`` stream so the values
are identical on every machine and numpy version — the committed golden outputs depend on it."""
import numpy as np
import torch

# the 5 real 10-D PCA blur-kernel codes of the reference's example data
# (Data/example_data/Set5/lr_random_blur/degradation_metadata.csv), rounded to 6 decimals
SET5_BLUR_CODES = np.array([
    [0.006943, 0.021855, 0.031484, 0.030229, 0.027254, 0.024536, 0.019063, 0.013387, 0.010976, 0.009122],
    [0.021203, 0.057692, 0.064342, 0.045704, 0.031924, 0.024084, 0.018079, 0.013024, 0.010881, 0.009122],
], dtype=np.float64)


def synth_state_dict(shapes, seed=8, zero_init_std=0.1):
    """shapes: ordered mapping key -> shape.  Weights/biases ~ U(-k, k), k = 1/sqrt(fan_in) (the scale
    of PyTorch's default conv init, so activations stay O(1) through 400 layers); 1-element params
    (the reference's zero-initialised `gamma`s) ~ N(0, zero_init_std) so those branches are exercised
    (SURVEY.md §7 hard part 6)."""
    rs = np.random.RandomState(seed)
    out = {}
    for key in sorted(shapes):
        shape = tuple(shapes[key])
        n = int(np.prod(shape)) if len(shape) else 1
        if n == 1:
            v = rs.normal(0.0, zero_init_std, size=shape)
        else:
            if key.endswith(".bias"):
                wshape = tuple(shapes[key[:-5] + ".weight"])
                fan_in = int(np.prod(wshape[1:]))
            else:
                fan_in = int(np.prod(shape[1:]))
            k = 1.0 / np.sqrt(fan_in)
            v = rs.uniform(-k, k, size=shape)
        out[key] = torch.from_numpy(np.asarray(v, dtype=np.float32).reshape(shape))
    return out


def synth_inputs(batch, height, width, num_metadata=10, seed=8):
    """LR images U[0,1) quantised to k/255 (as ToTensor would give) and metadata rows that mimic the
    10-D PCA blur codes (real codes lie in [0.003, 0.37])."""
    rs = np.random.RandomState(seed + 1000)
    x = np.floor(rs.uniform(0, 1, size=(batch, 3, height, width)) * 256.0) / 255.0
    meta = rs.uniform(0, 0.4, size=(batch, num_metadata))
    if num_metadata == 10:
        for i in range(min(batch, len(SET5_BLUR_CODES))):
            meta[i] = SET5_BLUR_CODES[i]
    x = torch.from_numpy(x.astype(np.float32))
    meta = torch.from_numpy(meta.astype(np.float32)).reshape(batch, num_metadata, 1, 1)
    return x, meta


def synth_target(shape, seed=8):
    """deterministic HR target for the training-step cases (L1 loss against it)"""
    rs = np.random.RandomState(seed + 2000)
    return torch.from_numpy(rs.uniform(0, 1, size=tuple(shape)).astype(np.float32))


N_PROJ = 8


def grad_projections(name, g):
    """dot products of a gradient with N_PROJ deterministic Gaussian directions (seeded by the parameter name):
    the compact fingerprint stored in tests/golden/grads_*.npz"""
    seed = int.from_bytes(name.encode()[-8:].rjust(8, b"\0"), "little") % (2 ** 31)
    rs = np.random.RandomState(seed)
    flat = g.detach().double().reshape(-1).numpy()
    return np.array([float(np.dot(rs.standard_normal(flat.size), flat)) for _ in range(N_PROJ)])
