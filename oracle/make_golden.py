"""Generates tests/golden/*.npz from the LIVE reference (TEST INFRASTRUCTURE ONLY).

Run in the build container, where the reference is mounted at /root/reference:

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.make_golden

Each case builds a reference network (imported through oracle/ref_shim.py), overwrites its
state_dict with the deterministic synthetic weights of oracle/synth.py, runs the reference's own
``forward`` on the deterministic inputs, and stores: the fp32 output, the parameter key -> shape
table (pins the checkpoint state_dict layout) and the case config.  Weights and inputs are NOT
stored — tests regenerate them from the same seeds.
"""
import json
import os
import sys

import numpy as np
import torch

from oracle.ref_shim import import_reference_architectures
from oracle.synth import synth_inputs, synth_state_dict

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# name -> (model, ctor kwargs, (B,H,W), num_metadata of the attributes tensor)
CASES = {
    # the published Q-RCAN configuration (q-rcan.toml: style standard + q layers), reduced depth
    "qrcan_standard_g2b2": ("qrcan", dict(n_resgroups=2, n_resblocks=2, style="standard", num_metadata=10,
                                          include_q_layer=True, scale=4), (2, 24, 20), 10),
    # full published depth (10 groups x 20 RCAB) on a small image
    "qrcan_standard_full": ("qrcan", dict(n_resgroups=10, n_resblocks=20, style="standard", num_metadata=10,
                                          include_q_layer=True, scale=4), (1, 32, 32), 10),
    "qrcan_noq_scale2": ("qrcan", dict(n_resgroups=1, n_resblocks=2, style="standard", num_metadata=10,
                                       include_q_layer=False, scale=2), (1, 12, 16), 10),
    "qrcan_modulate": ("qrcan", dict(n_resgroups=1, n_resblocks=2, style="modulate", num_metadata=1,
                                     include_q_layer=False, scale=4), (2, 12, 12), 64),
    "qrcan_max_concat_scale3": ("qrcan", dict(n_resgroups=1, n_resblocks=2, style="max_concat", num_metadata=10,
                                              include_q_layer=True, scale=3), (2, 12, 12), 10),
    "qrcan_softmax": ("qrcan", dict(n_resgroups=1, n_resblocks=2, style="softmax", num_metadata=10,
                                    include_q_layer=False, scale=4), (2, 12, 12), 10),
    "qrcan_mini_concat": ("qrcan", dict(n_resgroups=1, n_resblocks=2, style="mini_concat", num_metadata=10,
                                        include_q_layer=True, scale=4), (2, 12, 12), 10),
    "qrcan_extended_scale8": ("qrcan", dict(n_resgroups=1, n_resblocks=2, style="extended_attention",
                                            num_metadata=10, include_q_layer=True, scale=8), (1, 8, 8), 10),
    "qrcan_pa_selective": ("qrcan", dict(n_resgroups=3, n_resblocks=3, style="standard", num_metadata=11,
                                         include_q_layer=True, include_pixel_attention=True,
                                         selective_meta_blocks=[True, False, True],
                                         num_q_layers_inner_residual=2, scale=4), (1, 12, 12), 11),
    "qedsr_f64_b3": ("qedsr", dict(num_blocks=3, num_features=64, input_para=10, scale=4, res_scale=0.1,
                                   q_layer_nonlinearity=False), (2, 16, 12), 10),
    "qedsr_f256_b2_nl": ("qedsr", dict(num_blocks=2, num_features=256, input_para=10, scale=4, res_scale=0.1,
                                       q_layer_nonlinearity=True), (1, 12, 12), 10),
    "qsan_g2b2": ("qsan", dict(n_resgroups=2, n_resblocks=2, input_para=10, scale=4), (2, 16, 12), 10),
    "qhan_b1": ("qhan", dict(n_resgroups=10, n_resblocks=1, num_metadata=10, scale=4), (1, 8, 8), 10),
    # non-meta baselines (advanced/architectures.py): same kernels with the meta scale == 1
    "rcan_g2b2": ("rcan", dict(n_resgroups=2, n_resblocks=2, scale=4), (2, 20, 24), 1),
    "edsr_f64_b3": ("edsr", dict(num_blocks=3, net_features=64, scale=2, res_scale=0.1), (2, 12, 20), 1),
    "san_g2b2": ("san", dict(n_resgroups=2, n_resblocks=2, scale=4), (2, 16, 12), 1),
    "han_b1": ("han", dict(n_resgroups=10, n_resblocks=1, scale=4), (1, 8, 8), 1),
}


def build_reference(model, kwargs):
    arch = import_reference_architectures()
    cls = {"qrcan": arch.QRCAN, "qedsr": arch.QEDSR, "qsan": arch.QSAN, "qhan": arch.QHAN,
           "rcan": arch._ref_advanced.RCAN, "edsr": arch._ref_advanced.EDSR, "san": arch._ref_advanced.SAN,
           "han": arch._ref_advanced.HAN}[model]
    torch.manual_seed(8)
    return cls(**kwargs).eval()


def run_case(name):
    model, kwargs, (b, h, w), m_attr = CASES[name]
    net = build_reference(model, kwargs)
    shapes = {k: list(v.shape) for k, v in net.state_dict().items()}
    sd = synth_state_dict(shapes, seed=8)
    net.load_state_dict(sd, strict=True)
    x, meta = synth_inputs(b, h, w, num_metadata=m_attr, seed=8)
    with torch.no_grad():
        out = net(x) if model in ("rcan", "edsr", "san", "han") else net(x, meta)
    return shapes, out


def main(argv):
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    names = argv[1:] or list(CASES)
    for name in names:
        shapes, out = run_case(name)
        model, kwargs, bhw, m_attr = CASES[name]
        np.savez_compressed(
            os.path.join(GOLDEN_DIR, name + ".npz"),
            out=out.numpy().astype(np.float32),
            meta=np.frombuffer(json.dumps(dict(model=model, kwargs=kwargs, bhw=list(bhw), m_attr=m_attr,
                                               shapes=shapes, torch=torch.__version__)).encode(), dtype=np.uint8))
        print("%-28s out %s  mean|out| %.5f" % (name, tuple(out.shape), out.abs().mean().item()))


if __name__ == "__main__":
    main(sys.argv)
