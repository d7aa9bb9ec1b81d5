"""Generates tests/golden/grads_*.npz from the LIVE reference (TEST INFRASTRUCTURE ONLY).

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden_grads

For each case the reference network (weights from oracle/synth.py) runs forward on the deterministic inputs, the
reference's training criterion nn.L1Loss() (models/__init__.py:207-211) is taken against a deterministic target and
`loss.backward()` gives the parameter gradients exactly as `BaseModel.standard_update` (:481-489) would see them.
Stored per parameter: its L2 norm and its dot products with 8 deterministic random directions (float64) — enough
to pin a restatement of the backward pass without committing megabytes of gradients — plus the loss value.
"""
import json
import os
import sys

import numpy as np
import torch

from oracle.make_golden import CASES, GOLDEN_DIR, build_reference
from oracle.synth import N_PROJ, grad_projections, synth_inputs, synth_state_dict, synth_target

GRAD_CASES = ["qrcan_standard_g2b2", "qrcan_noq_scale2", "qrcan_modulate", "qrcan_max_concat_scale3", "qedsr_f64_b3",
              "qedsr_f256_b2_nl", "rcan_g2b2", "edsr_f64_b3", "qrcan_softmax", "qrcan_mini_concat", "qrcan_extended_scale8",
              # networks whose backward is not on the B200 path yet (tests/golden_util.py:NO_TRAINING_PATH): the
              # fingerprints pin the oracle's autograd for them ahead of the kernels
              "qrcan_pa_selective", "qsan_g2b2", "qhan_b1", "san_g2b2", "han_b1"]


def summarize(grads):
    names = sorted(grads)
    norms = np.array([float(grads[k].double().norm()) for k in names])
    proj = np.stack([grad_projections(k, grads[k]) for k in names])
    return names, norms, proj


def run_case(name):
    model, kwargs, (b, h, w), m_attr = CASES[name]
    net = build_reference(model, kwargs).train()
    shapes = {k: list(v.shape) for k, v in net.state_dict().items()}
    net.load_state_dict(synth_state_dict(shapes, seed=8), strict=True)
    x, meta = synth_inputs(b, h, w, num_metadata=m_attr, seed=8)
    out = net(x) if model in ("rcan", "edsr", "san", "han") else net(x, meta)
    y = synth_target(out.shape)
    loss = torch.nn.L1Loss()(out, y)
    net.zero_grad()
    loss.backward()
    grads = {k: p.grad for k, p in net.named_parameters() if p.grad is not None}
    return float(loss.detach()), grads


def main(argv):
    for name in (argv[1:] or GRAD_CASES):
        loss, grads = run_case(name)
        names, norms, proj = summarize(grads)
        np.savez_compressed(os.path.join(GOLDEN_DIR, "grads_" + name + ".npz"), norms=norms, proj=proj,
                            names=np.frombuffer(json.dumps(names).encode(), dtype=np.uint8), loss=np.float64(loss))
        print("%-28s loss %.6f  %d parameters, |g| in [%.3e, %.3e]" % (name, loss, len(names), norms.min(), norms.max()))


if __name__ == "__main__":
    main(sys.argv)
