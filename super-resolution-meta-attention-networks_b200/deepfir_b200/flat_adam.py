"""`torch.optim.Adam` whose `step()` is ONE kernel over flat buffers (`dfir_adam_step`).

The reference trains with `optim.Adam(params, lr)` (/root/reference/Code/SISR/models/__init__.py:291-299, stepped in
`standard_update` :481-489).  Q-RCAN has 1648 parameter tensors, most of them tiny: torch's fused implementation
still needs ~70 launches and 1.8 ms per step for 16 M parameters; one pass over flat storage is bandwidth bound
(28 B per parameter, ~0.1 ms).  This class stays a drop-in `torch.optim.Adam`:

* same constructor, `param_groups`, `zero_grad`, LR schedulers, `state_dict()` layout (per-parameter `step`,
  `exp_avg`, `exp_avg_sq` — the moment tensors are views of the flat buffers);
* at construction (and again whenever the parameters were moved) the parameters of the (single) group are re-homed
  into one flat fp32 buffer, 16-byte aligned per tensor, in `parameters()` order — the layout `deepfir_b200/train.py` uses for the flat gradient buffer, so the
  gradients are consumed where the backward wrote them;
* whenever that layout does not hold (several groups, gradients living elsewhere, amsgrad / maximize / non-CUDA or
  non-fp32 parameters, sparse gradients) `step()` simply defers to `torch.optim.Adam.step`.
"""
import ctypes as C

import torch

from . import _lib


def _aligned(n):
    return (n + 3) // 4 * 4


class FlatAdam(torch.optim.Adam):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, **kwargs):
        kwargs.pop("fused", None)
        kwargs.pop("foreach", None)
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, **kwargs)
        self._flat = None  # dict(params, offsets, total, p, m, v, step)
        if self._eligible():  # re-home the parameters now, before any kernel-format weight cache is built from them
            self._flatten()

    # ------------------------------------------------------------------ layout
    def _eligible(self):
        if len(self.param_groups) != 1:
            return False
        g = self.param_groups[0]
        if g.get("amsgrad") or g.get("maximize") or g.get("differentiable") or g.get("capturable"):
            return False
        ps = g["params"]
        return len(ps) > 0 and all(p.is_cuda and p.dtype == torch.float32 and p.device == ps[0].device for p in ps)

    def _homed(self):
        f = self._flat
        if f is None:
            return False
        ps = self.param_groups[0]["params"]
        if len(ps) != len(f["params"]) or any(a is not b for a, b in zip(ps, f["params"])):
            return False
        base = f["p"].data_ptr()  # parameters may have been moved by .to() / load_state_dict(assign=True): check both ends
        return ps[0].data_ptr() == base and ps[-1].data_ptr() == base + 4 * f["offsets"][-1]

    def _flatten(self):
        ps = self.param_groups[0]["params"]
        offsets, total = [], 0
        for p in ps:
            offsets.append(total)
            total += _aligned(p.numel())
        dev = ps[0].device
        old = self._flat
        flat_p = torch.zeros(total, device=dev, dtype=torch.float32)
        flat_m = torch.zeros(total, device=dev, dtype=torch.float32)
        flat_v = torch.zeros(total, device=dev, dtype=torch.float32)
        step = 0
        with torch.no_grad():
            for p, off in zip(ps, offsets):
                n = p.numel()
                flat_p[off:off + n].copy_(p.detach().reshape(-1))
                st = self.state.get(p)
                if st:  # moments loaded from a checkpoint or produced by earlier (unflattened) steps
                    flat_m[off:off + n].copy_(st["exp_avg"].reshape(-1))
                    flat_v[off:off + n].copy_(st["exp_avg_sq"].reshape(-1))
                    step = max(step, int(float(st["step"])))
                p.data = flat_p[off:off + n].view(p.shape)
        step_t = torch.tensor(float(step), dtype=torch.float32)
        for p, off in zip(ps, offsets):
            n = p.numel()
            self.state[p] = {"step": step_t, "exp_avg": flat_m[off:off + n].view(p.shape),
                             "exp_avg_sq": flat_v[off:off + n].view(p.shape)}
        self._flat = dict(params=list(ps), offsets=offsets, total=total, p=flat_p, m=flat_m, v=flat_v, step=step,
                          step_t=step_t)
        del old

    def _flat_grad(self):
        """the flat gradient buffer if every .grad is a view of ONE buffer with the parameters' own offsets"""
        f = self._flat
        ps = f["params"]
        g0 = ps[0].grad
        if g0 is None or not g0.is_cuda or g0.dtype != torch.float32:
            return None
        base = g0.data_ptr()
        offs = f["offsets"]
        for idx in range(1, len(ps)):  # every gradient must be the slice of the flat buffer at its parameter's offset
            g = ps[idx].grad
            if g is None or g.data_ptr() != base + 4 * offs[idx]:
                return None
        root = g0._base if g0._base is not None else g0
        if root.numel() < f["total"] or root.data_ptr() != base or not root.is_contiguous():
            return None
        return root

    # ------------------------------------------------------------------ torch.optim API
    @torch.no_grad()
    def step(self, closure=None):
        if not self._eligible():
            return super().step(closure)
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if not self._homed():
            self._flatten()
        gflat = self._flat_grad()
        if gflat is None:
            return self._fallback_step(loss)
        f = self._flat
        grp = self.param_groups[0]
        f["step"] += 1
        f["step_t"].fill_(float(f["step"]))
        lib = _lib.load_library()
        with torch.cuda.device(f["p"].device):
            rc = lib.dfir_adam_step(f["p"].data_ptr(), gflat.data_ptr(), f["m"].data_ptr(), f["v"].data_ptr(), f["total"],
                                    float(grp["lr"]), float(grp["betas"][0]), float(grp["betas"][1]), float(grp["eps"]),
                                    float(grp["weight_decay"]), f["step"],
                                    C.c_void_p(torch.cuda.current_stream(f["p"].device).cuda_stream))
        _lib.check(rc, "adam_step")
        # the kernel wrote through the flat alias, which no version counter saw: the networks' packed-weight cache
        # compares the tuple of parameter versions, so an in-place no-op on ONE parameter records the modification
        f["params"][0].add_(0)
        return loss

    def _fallback_step(self, loss):
        """gradients are not in the expected flat buffer: torch's own implementation on the same state tensors"""
        f = self._flat
        for st in self.state.values():  # per-parameter step tensors for the stock implementation
            st["step"] = torch.tensor(float(f["step"]), dtype=torch.float32)
        super().step(None)
        f["step"] += 1
        for st in self.state.values():
            st["step"] = f["step_t"]
        f["step_t"].fill_(float(f["step"]))
        return loss

    def state_dict(self):
        """torch.optim.Adam's layout with an INDEPENDENT `step` tensor per parameter: internally all parameters share
        one step tensor, and a checkpoint that kept the aliasing would make a stock Adam advance the step once per
        parameter and iteration after loading it (torch.save preserves storage sharing)."""
        sd = super().state_dict()
        for st in sd["state"].values():
            if "step" in st:
                st["step"] = torch.tensor(float(st["step"]), dtype=torch.float32)
        return sd

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._flat = None  # loaded moment tensors are fresh allocations: re-home on the next step
