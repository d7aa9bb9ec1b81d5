"""-m "not gpu": host-side logic — registry, state_dict layout, checkpoints, metadata formatting, the C ABI's
exported symbols, and the 2-rank sharding path on gloo."""
import ctypes
import os
import re
import subprocess
import sys
import tempfile

import pytest
import torch

from oracle import deepfir_oracle as O
from oracle.ref_shim import reference_available
from tests.golden_util import case_tensors, golden_names, load_golden, max_norm_err, oracle_forward

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
QRCAN_CASES = [n for n in golden_names() if n.startswith("qrcan")]


def test_registry_discovers_the_q_handlers():
    from SISR.models import available_models
    for name in ("qrcan", "qedsr", "qsan", "qhan"):
        assert available_models[name].endswith("handlers.%sHandler" % name.upper())


@pytest.mark.parametrize("name", QRCAN_CASES)
def test_state_dict_layout_matches_reference(name):
    """same parameter names, shapes AND order as the reference module (order matters: optimizer state dicts
    index parameters by position)."""
    from deepfir_b200.qrcan import QRCAN
    _, info = load_golden(name)
    sd = QRCAN(**info["kwargs"]).state_dict()
    assert list(sd.keys()) == list(info["shapes"].keys())
    assert all(list(v.shape) == info["shapes"][k] for k, v in sd.items())
    assert all(v.dtype == torch.float32 for v in sd.values())


@pytest.mark.parametrize("name", [n for n in golden_names() if n.startswith("qedsr")])
def test_qedsr_state_dict_layout_matches_reference(name):
    from deepfir_b200.qrcan import QEDSR
    _, info = load_golden(name)
    sd = QEDSR(**info["kwargs"]).state_dict()
    assert list(sd.keys()) == list(info["shapes"].keys())
    assert all(list(v.shape) == info["shapes"][k] for k, v in sd.items())


@pytest.mark.parametrize("name", ["qsan_g2b2", "qhan_b1"])
def test_qsan_qhan_state_dict_layout_matches_reference(name):
    from deepfir_b200.han_san import QHAN, QSAN
    _, info = load_golden(name)
    sd = (QSAN if info["model"] == "qsan" else QHAN)(**info["kwargs"]).state_dict()
    assert list(sd.keys()) == list(info["shapes"].keys())
    assert all(list(v.shape) == info["shapes"][k] for k, v in sd.items())


def test_forward_refuses_cpu_tensors_and_unsupported_options():
    from deepfir_b200.qrcan import QRCAN, ChannelAttentionParams
    net = QRCAN(n_resgroups=1, n_resblocks=1, style="standard", num_metadata=10)
    with pytest.raises(RuntimeError):
        net(torch.zeros(1, 3, 8, 8), torch.zeros(1, 10, 1, 1))
    with pytest.raises(RuntimeError):  # reference: 'Using an extreme channel attention reduction value'
        ChannelAttentionParams(64, "standard", reduction=8)
    with pytest.raises(RuntimeError):
        QRCAN(precision="int8")


def test_library_exports_every_declared_symbol():
    from deepfir_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "dfir.h")).read()
    declared = set(re.findall(r"\b(dfir_[a-z0-9_]+)\s*\(", hdr)) - {"dfir_qrcan_net"}
    if not os.path.isfile(_lib.lib_path()):  # fresh checkout: the .so is a build artefact (nvcc cross-compiles without a GPU)
        _lib.build_library()
    lib = ctypes.CDLL(_lib.lib_path())
    for sym in sorted(declared):
        assert hasattr(lib, sym), sym
    assert declared == set(_lib.PROTOTYPES), declared ^ set(_lib.PROTOTYPES)
    assert _lib.load_library().dfir_error_string(-5) == b"workspace too small"


def _handler(tmp, eval_mode=False, **kw):
    from SISR.models import ModelInterface
    return ModelInterface.define_model("qrcan", device=torch.device("cpu"), model_save_dir=tmp, eval_mode=eval_mode,
                                       lr=1e-4, scale=4, style="standard", metadata=["blur_kernel"],
                                       include_q_layer=True, n_resgroups=1, n_resblocks=2, **kw)


def test_handler_surface_and_checkpoint_roundtrip():
    with tempfile.TemporaryDirectory() as tmp:
        h = _handler(tmp, scheduler="cosine_annealing_warm_restarts",
                     scheduler_params=dict(t_mult=1, restart_period=100, lr_min=1e-7))
        assert (h.model_name, h.colorspace, h.im_input, h.num_metadata) == ("qrcan", "augmented_rgb", "unmodified", 10)
        assert h.legacy_load and h.style == "standard" and not h.channel_concat
        assert h.print_parameters() == sum(p.numel() for p in h.net.parameters())
        h.set_epoch(3)
        h.save_model("train_model", 3)
        state = torch.load(os.path.join(tmp, "train_model_3"), weights_only=False)
        assert set(state) == {"network", "optimizer", "model_name", "model_epoch", "scheduler_G"}
        h2 = _handler(tmp, scheduler="cosine_annealing_warm_restarts",
                      scheduler_params=dict(t_mult=1, restart_period=100, lr_min=1e-7))
        # legacy checkpoints carry 'model.module.' / 'model.' prefixes (reference :388-398)
        state["network"] = {"model.module." + k: v for k, v in state["network"].items()}
        h2.load_model("train_model", 3, legacy=True, preloaded_state=state)
        assert h2.curr_epoch == 3
        for (k1, v1), (k2, v2) in zip(h.net.state_dict().items(), h2.net.state_dict().items()):
            assert k1 == k2 and torch.equal(v1, v2)
        with pytest.raises(RuntimeError):  # eval-mode handlers cannot train (reference :467-468)
            _handler(tmp, eval_mode=True).run_train(torch.zeros(1, 3, 4, 4), torch.zeros(1, 3, 16, 16))


def test_generate_channels_matches_oracle_and_modulate_style():
    with tempfile.TemporaryDirectory() as tmp:
        h = _handler(tmp)
        md = torch.rand(3, 12, dtype=torch.float64)
        keys = [("qpi",) * 3] + [("blur_kernel",) * 3] * 10 + [("other",) * 3]
        got = h.generate_channels(torch.zeros(3, 3, 4, 4), md, keys)
        want = O.generate_channels(md, keys, ["blur_kernel"], 10)
        assert got.shape == (3, 10, 1, 1) and torch.equal(got, want)
        with pytest.raises(RuntimeError):
            h.generate_channels(torch.zeros(1, 3, 4, 4), None, keys)
        from SISR.models import ModelInterface
        hm = ModelInterface.define_model("qrcan", device=torch.device("cpu"), model_save_dir=tmp, eval_mode=True,
                                         n_resgroups=1, n_resblocks=1)  # defaults: style modulate, metadata qpi
        q = torch.tensor([[0.25], [0.75]], dtype=torch.float64)
        g = hm.generate_channels(torch.zeros(2, 3, 4, 4), q, [("qpi",) * 2])
        assert g.shape == (2, 64, 1, 1)
        assert torch.allclose(g, O.scale_qpi(q.float().reshape(2, 1, 1, 1)), rtol=1e-6)


def test_all_q_handlers_construct_on_cpu_and_refuse_cpu_forward():
    from SISR.models import ModelInterface
    for name, n_params in (("qedsr", None), ("qsan", 16353288), ("qhan", 16564545)):
        h = ModelInterface.define_model(name, device=torch.device("cpu"), model_save_dir="/tmp", eval_mode=True,
                                        metadata=["blur_kernel"])
        if n_params is not None:
            assert h.print_parameters() == n_params  # parameter counts of the reference
        with pytest.raises(RuntimeError):
            h.net(torch.zeros(1, 3, 8, 8), torch.zeros(1, 10, 1, 1))


@pytest.mark.skipif(not reference_available(), reason="live reference only exists in the build container")
def test_same_seed_same_initial_weights_and_checkpoint_interchange_with_live_reference():
    """construction order mirrors the reference, so torch.manual_seed(8) (the reference's default seed)
    yields identical initial weights; a state_dict moves in both directions with strict=True."""
    from oracle.ref_shim import import_reference_architectures
    from deepfir_b200.qrcan import QRCAN
    arch = import_reference_architectures()
    kw = dict(n_resgroups=2, n_resblocks=3, style="max_concat", num_metadata=10, include_q_layer=True,
              selective_meta_blocks=[True, False], num_q_layers_inner_residual=2, scale=4)
    torch.manual_seed(8)
    ref = arch.QRCAN(**kw)
    torch.manual_seed(8)
    ours = QRCAN(**kw)
    a, b = ref.state_dict(), ours.state_dict()
    assert list(a) == list(b)
    assert all(torch.equal(a[k], b[k]) for k in a)
    assert [n for n, _ in ref.named_parameters()] == [n for n, _ in ours.named_parameters()]
    ref.load_state_dict(ours.state_dict(), strict=True)
    ours.load_state_dict(ref.state_dict(), strict=True)


_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "super-resolution-meta-attention-networks_b200"))
from deepfir_b200.sharding import env_rank_world, run_sharded, max_over_ranks, shard_range
from oracle import deepfir_oracle as O
from tests.golden_util import load_golden, case_tensors, oracle_forward
rank, world, _ = env_rank_world()
dist.init_process_group("gloo", rank=rank, world_size=world)
ref, info = load_golden("qrcan_noq_scale2")
sd, x, meta = case_tensors(info)
x = torch.cat([x, x.flip(3), x.flip(2)], 0); meta = torch.cat([meta, meta * 0.5, meta * 2], 0)
with torch.no_grad():
    a, b, out = run_sharded(lambda xs, ms: oracle_forward(info, sd, xs, ms), x, meta, rank, world)
    full = oracle_forward(info, sd, x, meta)
assert (a, b) == shard_range(3, rank, world)
assert torch.allclose(out, full[a:b], atol=1e-6), "sharded != unsharded"
parts = [None] * world
dist.all_gather_object(parts, (a, b))
assert parts[0][1] == parts[1][0] and parts[0][0] == 0 and parts[-1][1] == 3
t = max_over_ranks(float(rank + 1))
assert t == float(world), t
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_two_rank_sharded_inference_on_gloo():
    """N>1 path on CPU: two ranks (gloo, 127.0.0.1) each run their slice of the image list (the oracle
    stands in for the GPU forward), slices tile the batch, results equal the unsharded run, and the timing
    reduction is a MAX over ranks."""
    with tempfile.NamedTemporaryFile("w", suffix=".py", delete=False) as fh:
        fh.write(_WORKER.format(root=ROOT))
        script = fh.name
    try:
        res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                              "--master-addr", "127.0.0.1", "--master-port", "29531", script],
                             capture_output=True, text=True, timeout=300,
                             env=dict(os.environ, OMP_NUM_THREADS="2", PYTHONDONTWRITEBYTECODE="1"))
    finally:
        os.unlink(script)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert res.stdout.count("ok") == 2


def test_shard_range_properties():
    from deepfir_b200.sharding import shard_range
    for n in (0, 1, 7, 32, 256):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


_DDP_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "super-resolution-meta-attention-networks_b200"))
from deepfir_b200.sharding import allreduce_mean_, env_rank_world, shard_range
from oracle.synth import synth_target
from tests.golden_util import load_golden, case_tensors, oracle_forward
rank, world, _ = env_rank_world()
dist.init_process_group("gloo", rank=rank, world_size=world)
ref, info = load_golden("qrcan_standard_g2b2")
sd, x, meta = case_tensors(info)            # 2 images: one per rank
y = synth_target(ref.shape)


def grads(xs, ms, ys):
    leaves = {{k: v.clone().requires_grad_(True) for k, v in sd.items()}}
    torch.nn.functional.l1_loss(oracle_forward(info, leaves, xs, ms), ys).backward()
    return torch.cat([leaves[k].grad.reshape(-1) for k in sorted(leaves)])


a, b = shard_range(x.shape[0], rank, world)
flat = grads(x[a:b], meta[a:b], y[a:b])      # this rank's gradient of ITS mean-L1 loss, one flat buffer
allreduce_mean_(flat)                        # the data-parallel exchange step
full = grads(x, meta, y)                     # what one process sees on the whole batch
err = float((flat - full).norm() / full.norm())
assert err < 1e-5, err
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_two_rank_gradient_allreduce_equals_full_batch_gradient_on_gloo():
    """training N>1 path on CPU (gloo, world 2): per-rank mean-L1 gradients, averaged over ranks through the flat
    buffer all-reduce the GPU path uses (deepfir_b200/train.py), equal the full-batch gradient."""
    with tempfile.NamedTemporaryFile("w", suffix=".py", delete=False) as fh:
        fh.write(_DDP_WORKER.format(root=ROOT))
        script = fh.name
    try:
        res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                              "--master-addr", "127.0.0.1", "--master-port", "29533", script],
                             capture_output=True, text=True, timeout=300,
                             env=dict(os.environ, OMP_NUM_THREADS="2", PYTHONDONTWRITEBYTECODE="1"))
    finally:
        os.unlink(script)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert res.stdout.count("ok") == 2


def test_bench_clock_sampler_uses_only_lines_inside_the_timed_region(tmp_path):
    """bench.py starts nvidia-smi ahead of the warm-up; only the lines stamped inside [mark_begin, mark_end] may reach
    the `clocks` object, throttle reasons included (a warm-up line must not leak a reason into the timed region)."""
    import datetime
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)

    class _Done:
        def terminate(self):
            pass

        def wait(self, timeout=None):
            return 0

    def stamp(t):
        return datetime.datetime.fromtimestamp(t).strftime("%Y/%m/%d %H:%M:%S.%f")[:-3]

    t0 = 1_792_000_000.0
    lines = [(t0 - 0.30, 1200, "Active", "Not Active"),      # warm-up: ignored
             (t0 + 0.02, 1950, "Not Active", "Not Active"),
             (t0 + 0.04, 1905, "Not Active", "Active"),
             (t0 + 0.06, 1965, "Not Active", "Not Active"),
             (t0 + 0.50, 1000, "Active", "Active")]           # after the timed steps: ignored
    path = tmp_path / "smi.csv"
    path.write_text("".join("%s, %d, 1965, %s, Not Active, Not Active, %s\n" % (stamp(t), mhz, hw, cap)
                            for t, mhz, hw, cap in lines) + "garbage line\n")
    s = bench.ClockSampler(0)
    s.proc, s.path, s.t0, s.t1 = _Done(), str(path), t0, t0 + 0.1
    out = s.stop()
    assert out["samples"] == 3 and out["sm_mhz"] == 1950.0 and out["sm_max_mhz"] == 1965.0
    assert out["reasons"] == ["sw_power_cap"] and out["window"] == "device-timed steps"
    # nothing inside the window: falls back to the window that also covers the end-to-end steps, else says so
    path.write_text("%s, 1800, 1965, Not Active, Not Active, Not Active, Not Active\n" % stamp(t0 + 0.3))
    s = bench.ClockSampler(0)
    s.proc, s.path, s.t0, s.t1, s.t1_fallback = _Done(), str(path), t0, t0 + 0.1, t0 + 0.4
    out = s.stop()
    assert out["samples"] == 1 and "end-to-end" in out["window"]
    path.write_text("")
    s = bench.ClockSampler(0)
    s.proc, s.path, s.t0, s.t1 = _Done(), str(path), t0, t0 + 0.1
    out = s.stop()
    assert out["samples"] == 0 and out["sm_mhz"] is None


def test_set_multi_gpu_refuses_data_parallel_replication():
    """the reference wraps the net in nn.DataParallel (models/__init__.py:307); the B200 networks hold per-device pointer
    tables and must not be replicated that way: more than one device raises and points at one process per GPU"""
    import tempfile
    from SISR.models import ModelInterface
    h = ModelInterface.define_model("qrcan", device=torch.device("cpu"), model_save_dir=tempfile.gettempdir(), eval_mode=True,
                                    metadata=["blur_kernel"], n_resgroups=1, n_resblocks=1)
    with pytest.raises(NotImplementedError, match="one process per GPU"):
        h.set_multi_gpu(device_ids=[0, 1])
    h.set_multi_gpu(device_ids=[])          # nothing to do: no wrapper module appears
    assert not isinstance(h.net, torch.nn.DataParallel)


def test_unsupported_qedsr_width_fails_loudly():
    from deepfir_b200.qrcan import QEDSR
    with pytest.raises(RuntimeError, match="not supported in fp32 mode"):
        QEDSR(num_features=192, num_blocks=1, input_para=10, precision="fp32")
    QEDSR(num_features=192, num_blocks=1, input_para=10)  # bf16 inference: three 64-channel planes
    QEDSR(num_features=128, num_blocks=1, input_para=10, precision="fp32")


def test_flat_adam_state_dict_has_one_step_tensor_per_parameter():
    """the checkpoint must be loadable by a stock torch.optim.Adam: independent step tensors"""
    from deepfir_b200.flat_adam import FlatAdam
    ps = [torch.nn.Parameter(torch.randn(3, 3)) for _ in range(4)]
    opt = FlatAdam(ps, lr=1e-3)
    sum((p ** 2).sum() for p in ps).backward()
    opt.step()
    sd = opt.state_dict()
    steps = [st["step"] for st in sd["state"].values()]
    assert len(steps) == 4 and all(float(t) == 1.0 for t in steps)
    assert len({id(t) for t in steps}) == 4 and len({t.untyped_storage().data_ptr() for t in steps}) == 4


_CHOP_WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "super-resolution-meta-attention-networks_b200"))
from deepfir_b200.sharding import env_rank_world, quadrant_geometry, run_chopped_sharded
from tests.golden_util import load_golden, case_tensors, oracle_forward
rank, world, _ = env_rank_world()
dist.init_process_group("gloo", rank=rank, world_size=world)
ref, info = load_golden("qrcan_noq_scale2")      # scale 2, 1 group x 2 blocks: cheap on the CPU
sd, _, _ = case_tensors(info)
g = torch.Generator().manual_seed(4)
x = torch.rand(3, 3, 25, 30, generator=g)          # odd height: unequal kept parts
meta = torch.rand(3, 10, 1, 1, generator=g) * 0.4
fwd = lambda xs, ms: oracle_forward(info, sd, xs, ms)
with torch.no_grad():
    sharded = run_chopped_sharded(fwd, x, meta, 2, rank, world, shave=10, out_channels=3)
    single = run_chopped_sharded(fwd, x, meta, 2, 0, 1, shave=10)
    # the single-process result is the reference's quadrant loop
    want = torch.empty_like(single)
    for lr_r, lr_c, dr, dc, sr_r, sr_c in quadrant_geometry(25, 30, 2, 10):
        want[:, :, dr, dc] = fwd(x[:, :, lr_r, lr_c].contiguous(), meta)[:, :, sr_r, sr_c]
assert torch.equal(single, want)
assert float((sharded - single).abs().max()) <= 1e-6, float((sharded - single).abs().max())
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_two_rank_tile_sharding_of_one_image_on_gloo():
    """inference sharded by LR tile (north_star): the reference's four overlapping quadrants of a batch spread over two
    ranks, stitched with one all-reduce, equal the single-process chop (handlers.py:99-137)"""
    with tempfile.NamedTemporaryFile("w", suffix=".py", delete=False) as fh:
        fh.write(_CHOP_WORKER.format(root=ROOT))
        script = fh.name
    try:
        res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                              "--master-addr", "127.0.0.1", "--master-port", "29535", script],
                             capture_output=True, text=True, timeout=300,
                             env=dict(os.environ, OMP_NUM_THREADS="2", PYTHONDONTWRITEBYTECODE="1"))
    finally:
        os.unlink(script)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert res.stdout.count("ok") == 2
