"""BASELINE.json configs[2]: Q-EDSR x4 (32 resblocks, 256 ch) inference on synthetic 480x270 LR frames, one GPU."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "super-resolution-meta-attention-networks_b200"))
import torch
from deepfir_b200.qrcan import QEDSR
torch.manual_seed(8)
kw = dict(num_blocks=32, num_features=256, input_para=10, scale=4, res_scale=0.1, q_layer_nonlinearity=False)
g = torch.Generator().manual_seed(8)
x = torch.rand(1, 3, 270, 480, generator=g).cuda()
meta = (torch.rand(1, 10, 1, 1, generator=g) * 0.4).cuda()
FLOP = 100505088 * 270 * 480
def timeit(fn, n):
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
net = QEDSR(precision="bf16", **kw).cuda().eval()
with torch.no_grad():
    for _ in range(2): out = net(x, meta)
    ms = timeit(lambda: net(x, meta), 5)
    print("bf16 tensor-core (64-channel planes), eager launches: %.2f ms/frame = %.1f fps, %.1f out MPix/s, %.0f TFLOP/s"
          % (ms, 1e3 / ms, 1080 * 1920 / 1e6 / (ms / 1e3), FLOP / ms / 1e9))
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        outg = net(x, meta)
    graph.replay()
    ms = timeit(graph.replay, 5)
    print("same, replayed from a CUDA graph (%d launches):       %.2f ms/frame = %.1f fps, %.1f out MPix/s, %.0f TFLOP/s"
          % (net._wide.launch_count(), ms, 1e3 / ms, 1080 * 1920 / 1e6 / (ms / 1e3), FLOP / ms / 1e9))
    print("graph vs eager max diff", float((outg - out).abs().max()))
if len(sys.argv) > 1 and sys.argv[1] == "fp32":
    net32 = QEDSR(precision="fp32", **kw).cuda().eval()
    net32.load_state_dict(net.state_dict())
    with torch.no_grad():
        o32 = net32(x, meta)
        ms = timeit(lambda: net32(x, meta), 1)
    print("fp32 CUDA-core path: %.1f ms/frame = %.2f fps; max|bf16-fp32|/max|fp32| = %.2e"
          % (ms, 1e3 / ms, float((out - o32).abs().max() / o32.abs().max())))
