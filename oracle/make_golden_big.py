"""Generates tests/golden/big_*.npz — fingerprints of the LIVE reference at the BASELINE.json shapes (TEST
INFRASTRUCTURE ONLY).

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden_big [case ...]

The outputs at these shapes are too large to commit (C3: 1x3x1080x1920 fp32 = 25 MB), so each fixture keeps
  * `sub`    : the reference output sampled every `stride` pixels in y and x (fp32),
  * `proj`   : dot products of the full output with 8 seeded Gaussian directions (float64), its L2 norm and max |.|,
  * `meta`   : the case description and the parameter key -> shape table.
tests/test_big_shapes_gpu.py compares the B200 output with `sub` (every sampled pixel) and with the projections, and
pins the oracle at the same shape with the same fingerprint before using its full output for the PSNR bar.

Cases (BASELINE.json configs[...]):
  C2 big_qrcan_full_2x128   Q-RCAN x4, 10 groups x 20 RCAB, 2 x 128x128 LR images
  C3 big_qedsr_f256_270x480 Q-EDSR x4, 256 features (2 blocks), one 480x270 LR frame (ragged 480 = 3*128 + 96)
  C5 big_qsan_g2b2_128      Q-SAN x4 (2 groups x 2 blocks, SOCA + non-local), one 128x128 LR image
  C4 grads_big_qrcan_full_16x64  (make_golden_grads-style gradient fingerprint) full depth, 16 x 64x64 patches
"""
import json
import os
import sys
import time

import numpy as np
import torch

from oracle.make_golden import GOLDEN_DIR, build_reference
from oracle.synth import synth_inputs, synth_state_dict, synth_target

BIG_CASES = {
    "big_qrcan_full_2x128": ("qrcan", dict(n_resgroups=10, n_resblocks=20, style="standard", num_metadata=10,
                                           include_q_layer=True, scale=4), (2, 128, 128), 10, 8),
    "big_qedsr_f256_270x480": ("qedsr", dict(num_blocks=2, num_features=256, input_para=10, scale=4, res_scale=0.1,
                                             q_layer_nonlinearity=False), (1, 270, 480), 10, 12),
    "big_qsan_g2b2_128": ("qsan", dict(n_resgroups=2, n_resblocks=2, input_para=10, scale=4), (1, 128, 128), 10, 8),
}
BIG_GRAD_CASES = {
    "big_qrcan_full_16x64": ("qrcan", dict(n_resgroups=10, n_resblocks=20, style="standard", num_metadata=10,
                                           include_q_layer=True, scale=4), (16, 64, 64), 10),
}
N_OUT_PROJ = 8


def output_fingerprint(out, stride):
    """(sub, proj, norm, amax) of an NCHW output tensor; directions are seeded by the output shape only"""
    o = out.detach().double().numpy()
    sub = np.ascontiguousarray(out.detach().numpy()[:, :, ::stride, ::stride]).astype(np.float32)
    rs = np.random.RandomState(4242)
    flat = o.reshape(-1)
    proj = np.empty(N_OUT_PROJ)
    for i in range(N_OUT_PROJ):  # one direction at a time: the C3 output has 6.2 M elements
        proj[i] = float(np.dot(rs.standard_normal(flat.size), flat))
    return sub, proj, float(np.linalg.norm(flat)), float(np.abs(flat).max())


def run_forward(name):
    model, kwargs, (b, h, w), m_attr, stride = BIG_CASES[name]
    net = build_reference(model, kwargs)
    shapes = {k: list(v.shape) for k, v in net.state_dict().items()}
    net.load_state_dict(synth_state_dict(shapes, seed=8), strict=True)
    x, meta = synth_inputs(b, h, w, num_metadata=m_attr, seed=8)
    with torch.no_grad():
        out = net(x, meta)
    return shapes, out


def run_grads(name):
    from oracle.make_golden_grads import summarize
    model, kwargs, (b, h, w), m_attr = BIG_GRAD_CASES[name]
    net = build_reference(model, kwargs).train()
    shapes = {k: list(v.shape) for k, v in net.state_dict().items()}
    net.load_state_dict(synth_state_dict(shapes, seed=8), strict=True)
    x, meta = synth_inputs(b, h, w, num_metadata=m_attr, seed=8)
    out = net(x, meta)
    loss = torch.nn.L1Loss()(out, synth_target(out.shape))
    net.zero_grad()
    loss.backward()
    grads = {k: p.grad for k, p in net.named_parameters() if p.grad is not None}
    names, norms, proj = summarize(grads)
    return shapes, float(loss.detach()), names, norms, proj


def main(argv):
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    names = argv[1:] or (list(BIG_CASES) + ["grads_" + n for n in BIG_GRAD_CASES])
    for name in names:
        t0 = time.time()
        if name.startswith("grads_"):
            model, kwargs, bhw, m_attr = BIG_GRAD_CASES[name[6:]]
            shapes, loss, pnames, norms, proj = run_grads(name[6:])
            np.savez_compressed(
                os.path.join(GOLDEN_DIR, name + ".npz"), norms=norms, proj=proj,
                names=np.frombuffer(json.dumps(pnames).encode(), dtype=np.uint8), loss=np.float64(loss),
                meta=np.frombuffer(json.dumps(dict(model=model, kwargs=kwargs, bhw=list(bhw), m_attr=m_attr,
                                                   shapes=shapes, torch=torch.__version__)).encode(), dtype=np.uint8))
            print("%-32s loss %.6f  %d parameters  (%.0f s)" % (name, loss, len(pnames), time.time() - t0))
            continue
        model, kwargs, bhw, m_attr, stride = BIG_CASES[name]
        shapes, out = run_forward(name)
        sub, proj, norm, amax = output_fingerprint(out, stride)
        np.savez_compressed(
            os.path.join(GOLDEN_DIR, name + ".npz"), sub=sub, proj=proj, norm=np.float64(norm), amax=np.float64(amax),
            meta=np.frombuffer(json.dumps(dict(model=model, kwargs=kwargs, bhw=list(bhw), m_attr=m_attr, stride=stride,
                                               out_shape=list(out.shape), shapes=shapes,
                                               torch=torch.__version__)).encode(), dtype=np.uint8))
        print("%-32s out %s  |out| %.4f  max %.4f  (%.0f s)" % (name, tuple(out.shape), norm, amax, time.time() - t0))


if __name__ == "__main__":
    main(sys.argv)
