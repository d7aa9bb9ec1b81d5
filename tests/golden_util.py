import json
import os

import numpy as np
import torch

from oracle import deepfir_oracle as O
from oracle.synth import synth_inputs, synth_state_dict

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names():
    """full-output fixtures (small shapes); the BASELINE-shape fingerprints big_*.npz have their own loader"""
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR)
                  if f.endswith(".npz") and not f.startswith("grads_") and not f.startswith("big_"))


def grad_golden_names():
    return sorted(f[6:-4] for f in os.listdir(GOLDEN_DIR)
                  if f.endswith(".npz") and f.startswith("grads_") and not f.startswith("grads_big_"))


def big_golden_names():
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz") and f.startswith("big_"))


def load_big_golden(name):
    """fingerprint of the reference output at a BASELINE.json shape (oracle/make_golden_big.py): the output sampled every
    `stride` pixels, 8 seeded projections of the full output, its norm and max"""
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    info = json.loads(bytes(z["meta"]).decode())
    return dict(sub=torch.from_numpy(z["sub"]), proj=z["proj"], norm=float(z["norm"]), amax=float(z["amax"])), info


def big_fingerprint_errors(out, fp, info):
    """(max |out - ref| / max |ref| over the sampled pixels, worst |projection error| / ||ref||, PSNR over the sampled
    pixels with peak 1.0)"""
    st = info["stride"]
    sub = out[:, :, ::st, ::st].double()
    ref = fp["sub"].double()
    err_sub = float((sub - ref).abs().max() / fp["amax"])
    rs = np.random.RandomState(4242)
    flat = out.detach().double().reshape(-1).numpy()
    worst = 0.0
    for i in range(len(fp["proj"])):
        worst = max(worst, abs(float(np.dot(rs.standard_normal(flat.size), flat)) - fp["proj"][i]) / fp["norm"])
    mse = float(((sub - ref) ** 2).mean())
    psnr = 10.0 * np.log10(1.0 / max(mse, 1e-30))
    return err_sub, worst, psnr


def load_big_grad_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, "grads_" + name + ".npz"))
    names = json.loads(bytes(z["names"]).decode())
    info = json.loads(bytes(z["meta"]).decode())
    return float(z["loss"]), {k: (float(z["norms"][i]), z["proj"][i]) for i, k in enumerate(names)}, info


# networks that have gradient fingerprints (the oracle's backward is pinned for them) but no training path in the B200
# library: none since round 2 (pixel attention, Q-SAN / SAN and Q-HAN / HAN train through the library).
NO_TRAINING_PATH = ()


def trainable_grad_golden_names():
    return [n for n in grad_golden_names() if n not in NO_TRAINING_PATH]


def load_grad_golden(name):
    """(loss, {parameter name: (norm, projections[8])}) of the reference's own backward pass"""
    z = np.load(os.path.join(GOLDEN_DIR, "grads_" + name + ".npz"))
    names = json.loads(bytes(z["names"]).decode())
    return float(z["loss"]), {k: (float(z["norms"][i]), z["proj"][i]) for i, k in enumerate(names)}


def oracle_grads(info, sd, x, meta):
    """loss and parameter gradients of the oracle under the reference's training criterion (L1, mean) against the
    deterministic target: autograd through the CPU restatement"""
    from oracle.synth import synth_target
    leaves = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}
    out = oracle_forward(info, leaves, x, meta)
    y = synth_target(out.shape)
    loss = torch.nn.functional.l1_loss(out, y)
    loss.backward()
    return float(loss.detach()), out.detach(), y, {k: v.grad for k, v in leaves.items() if v.grad is not None}


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    info = json.loads(bytes(z["meta"]).decode())
    return torch.from_numpy(z["out"]), info


def case_tensors(info, seed=8):
    sd = synth_state_dict(info["shapes"], seed=seed)
    b, h, w = info["bhw"]
    x, meta = synth_inputs(b, h, w, num_metadata=info["m_attr"], seed=seed)
    return sd, x, meta


def oracle_forward(info, sd, x, meta, nm=O.EXACT):
    model, kw = info["model"], info["kwargs"]
    if model == "qrcan":
        return O.qrcan_forward(x, meta, sd, style=kw.get("style", "modulate"), nm=nm)
    if model == "qedsr":
        return O.qedsr_forward(x, meta, sd, res_scale=kw.get("res_scale", 0.1), nm=nm)
    if model == "qsan":
        return O.qsan_forward(x, meta, sd, nm=nm)
    if model == "qhan":
        return O.qhan_forward(x, meta, sd, nm=nm)
    if model == "rcan":
        return O.rcan_forward(x, sd, nm=nm)
    if model == "san":
        return O.san_forward(x, sd, nm=nm)
    if model == "han":
        return O.qhan_forward(x, None, sd, nm=nm)
    if model == "edsr":
        return O.edsr_forward(x, sd, res_scale=kw.get("res_scale", 0.1), nm=nm)
    raise KeyError(model)


def max_norm_err(a, b):
    """SURVEY.md §8d fp32-mode metric: max|a-b| / max|b|."""
    return ((a.double() - b.double()).abs().max() / b.double().abs().max()).item()
