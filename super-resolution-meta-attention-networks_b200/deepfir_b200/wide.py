"""Q-EDSR with more than 64 features (the published 256-feature configuration, BASELINE.json configs[2]) on the
tensor cores.

Reference: QEDSR / ParamResBlock (/root/reference/Code/SISR/models/attention_manipulators/architectures.py:332-399).
Every activation is kept as C/64 planes of 64 channels (NHWC bf16 for the conv operands, fp32 for the residual stream),
so a C -> C convolution is a (C/64) x (C/64) block matrix of the 64 -> 64 tcgen05 convolution; the sums over input
chunks are accumulated by chaining launches through the kernel's skip input — in the trunk on the hi / 8-bit lo stream format
(`dfir_conv3x3_c64_accumulate_hl8`: 24-bit floats, 8 instead of 14 B per element and the TMA-tile epilogue), in the upsampler in
fp32 (`dfir_conv3x3_c64_accumulate`).  ParamResBlock's `conv2(.) * res_scale * meta + x` is accumulated in place on the stream
planes, the upsampler's PixelShuffle is folded into the TMA store of the last chunk, the 256 -> 3 tail accumulates into the NCHW
output.
No weights or activations take a detour through the CPU; all launches go to the caller's stream.
"""
import ctypes as C

import torch

from . import _lib


def _st(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


class WideQEDSR:
    """kernel-format weights + forward of a QEDSR whose feature width is a multiple of 64 (> 64)"""

    def __init__(self, net):
        lib = _lib.load_library()
        cfg = net.cfg
        self.F = F = cfg["n_feats"]
        assert F % 64 == 0 and F > 64
        self.nc = nc = F // 64
        self.scale = net.scale
        self.r = 3 if net.scale == 3 else 2
        self.M = cfg["num_metadata"]
        self.hid = cfg["meta_hidden"]
        self.res_scale = float(cfg["res_scale"])
        self.meta_relu = int(cfg["meta_relu"])
        self.out_feats = cfg["out_feats"]
        self.in_feats = cfg["in_feats"]
        dev = net.head.weight.device
        self.dev = dev
        self.keep = []
        with torch.no_grad(), torch.cuda.device(dev):
            st = _st(dev)

            def pack(w64, nt_rows=64):  # OIHW [<=64][64][3][3] fp32 -> one swizzled tile
                w64 = w64.contiguous().float()
                out = torch.empty(9 * nt_rows * 128, dtype=torch.uint8, device=dev)
                _lib.check(lib.dfir_pack_conv3x3_bf16(w64.data_ptr(), out.data_ptr(), w64.shape[0], 64, nt_rows, 0, 1, st),
                           "pack wide tile")
                self.keep.append(out)
                return out

            def tiles(conv):  # [j][i] tiles of a C -> C conv, biases per output chunk
                w = conv.weight
                return ([[pack(w[j * 64:(j + 1) * 64, i * 64:(i + 1) * 64]) for i in range(nc)] for j in range(nc)],
                        [conv.bias[j * 64:(j + 1) * 64].contiguous().float() for j in range(nc)])

            # head 3 -> F: CUDA-core kernel, one launch per output chunk ([9][Cin][64] fp32 weights)
            hw = net.head.weight.float()
            self.head_w, self.head_b = [], []
            for j in range(nc):
                wj = hw[j * 64:(j + 1) * 64].contiguous()
                out = torch.empty(9 * self.in_feats * 64, dtype=torch.float32, device=dev)
                _lib.check(lib.dfir_pack_conv3x3_f32(wj.data_ptr(), out.data_ptr(), 64, self.in_feats, st), "pack head")
                self.head_w.append(out)
                self.head_b.append(net.head.bias[j * 64:(j + 1) * 64].contiguous().float())
            self.blocks = [(tiles(blk.body[0]), tiles(blk.body[2])) for blk in net.body]
            self.final = tiles(net.final_body)
            # upsampler: conv F -> r*r*F + PixelShuffle(r); output channel (c, s) = c*r*r + s
            rr = self.r * self.r
            self.ups = []
            for conv in [m for m in net.tail[0] if isinstance(m, torch.nn.Conv2d)]:
                w = conv.weight.view(F, rr, F, 3, 3)
                b = conv.bias.view(F, rr)
                self.ups.append(([[[pack(w[jc * 64:(jc + 1) * 64, s, i * 64:(i + 1) * 64]) for i in range(nc)]
                                   for s in range(rr)] for jc in range(nc)],
                                 [[b[jc * 64:(jc + 1) * 64, s].contiguous().float() for s in range(rr)] for jc in range(nc)]))
            tw = net.tail[1].weight
            self.tail_w = [pack(tw[:, i * 64:(i + 1) * 64], nt_rows=16) for i in range(nc)]
            tb = torch.zeros(16, device=dev)
            tb[: self.out_feats] = net.tail[1].bias
            self.tail_b = tb
            # meta-attention MLPs, one parameter set per output chunk
            nblk = len(net.body)
            hid, M = self.hid, self.M
            z = lambda *shape: torch.zeros(*shape, device=dev, dtype=torch.float32)
            self.meta = []
            for j in range(nc):
                w1, b1, w2, b2 = z(nblk, hid, M), z(nblk, hid), z(nblk, 64, hid), z(nblk, 64)
                for k, blk in enumerate(net.body):
                    f1, f2 = blk.attention_layer.fcs()
                    w1[k], b1[k] = f1.weight.reshape(hid, M), f1.bias
                    w2[k], b2[k] = f2.weight.reshape(F, hid)[j * 64:(j + 1) * 64], f2.bias[j * 64:(j + 1) * 64]
                self.meta.append((w1, b1, w2, b2))
        self._buf = {}
        self._graphs = {}

    # -------------------------------------------------------------------------------------------------
    def _buffers(self, B, H, W):
        key = (B, H, W)
        b = self._buf.get(key)
        if b is None:
            self._buf.clear()
            dev, nc, r = self.dev, self.nc, self.r
            f32 = lambda *s: torch.empty(*s, device=dev, dtype=torch.float32)
            bf = lambda *s: torch.empty(*s, device=dev, dtype=torch.bfloat16)
            i8 = lambda *s: torch.empty(*s, device=dev, dtype=torch.int8)
            # residual stream and partial sums in the hi / 8-bit lo format (24-bit floats: DESIGN.md section 3): H = head output,
            # X = stream, P = partial sums of a K-chunk chain, Z = zeros (skip of a chain's first chunk)
            b = dict(Hs=[f32(B, H, W, 64) for _ in range(nc)],
                     Hhi=[bf(B, H, W, 64) for _ in range(nc)], Hlo=[i8(B, H, W, 64) for _ in range(nc)],
                     Xhi=[bf(B, H, W, 64) for _ in range(nc)], Xlo=[i8(B, H, W, 64) for _ in range(nc)],
                     Phi=bf(B, H, W, 64), Plo=i8(B, H, W, 64),
                     Zhi=torch.zeros(B, H, W, 64, device=dev, dtype=torch.bfloat16),
                     Zlo=torch.zeros(B, H, W, 64, device=dev, dtype=torch.int8),
                     T=[bf(B, H, W, 64) for _ in range(nc)], Fbf=[bf(B, H, W, 64) for _ in range(nc)],
                     sq=[f32(len(self.blocks), B, 64) for _ in range(nc)], U=[])
            h, w = H, W
            for _ in self.ups:
                h, w = h * r, w * r
                b["U"].append([bf(B, h, w, 64) for _ in range(nc)])
            big = B * h * w * 64 // (r * r)  # partial sums live at the INPUT resolution of the last stage
            b["P"] = f32(max(B * H * W * 64, big))
            b["junk"] = bf(max(B * H * W * 64, big))
            self._buf[key] = b
        return b

    def forward(self, x, attr):
        """One frame batch through the plane loops.  A forward is ~1200 launches of 12-25 us: issued from Python they are
        host-bound (28.7 ms per 270x480 frame), so from the second call of a shape on the launches are replayed from a CUDA
        graph (static input / output buffers; `cuda_graphs = False` on the object turns it off)."""
        if not getattr(self, "cuda_graphs", True) or torch.cuda.is_current_stream_capturing():
            return self._forward(x, attr)
        key = (tuple(x.shape), tuple(attr.shape))
        g = self._graphs.get(key)
        if g is None:  # first call of this shape: eager (also sets the kernels' attributes, which capture must not do)
            self._graphs.clear()
            self._graphs[key] = dict(graph=None)
            return self._forward(x, attr)
        with torch.cuda.device(x.device):
            if g["graph"] is None:
                g["x"], g["attr"] = torch.empty_like(x), torch.empty_like(attr)
                g["x"].copy_(x)
                g["attr"].copy_(attr)
                torch.cuda.synchronize(x.device)
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    g["out"] = self._forward(g["x"], g["attr"])
                g["graph"] = graph
            g["x"].copy_(x)
            g["attr"].copy_(attr)
            g["graph"].replay()
            return g["out"].clone()

    def _forward(self, x, attr):
        lib = _lib.load_library()
        nc, r = self.nc, self.r
        B, _, H, W = x.shape
        dev = x.device
        st = _st(dev)
        bufs = self._buffers(B, H, W)
        Hs, Hhi, Hlo, Xhi, Xlo, T, Fbf, P, junk = (bufs[k] for k in ("Hs", "Hhi", "Hlo", "Xhi", "Xlo", "T", "Fbf", "P", "junk"))
        Phi, Plo, Zhi, Zlo = bufs["Phi"], bufs["Plo"], bufs["Zhi"], bufs["Zlo"]
        ptr = lambda t: None if t is None else t.data_ptr()

        def acc(inp, w, bias, svec, skip, out32, outbf, relu, h, wd):   # fp32 running sum (upsampler stages)
            _lib.check(lib.dfir_conv3x3_c64_accumulate(inp.data_ptr(), w.data_ptr(), ptr(bias), B, h, wd, ptr(svec), ptr(skip),
                                                       ptr(out32), outbf.data_ptr(), relu, st), "wide conv")

        def acc8(inp, w, bias, svec, skip, out, relu):   # running sum in the hi / 8-bit lo format; skip, out = (hi, lo8)
            _lib.check(lib.dfir_conv3x3_c64_accumulate_hl8(inp.data_ptr(), w.data_ptr(), ptr(bias), B, H, W, ptr(svec),
                                                           skip[0].data_ptr(), skip[1].data_ptr(), out[0].data_ptr(),
                                                           ptr(out[1]), relu, st), "wide conv hl8")

        nblk = len(self.blocks)
        npl = B * H * W * 64
        for j in range(nc):
            _lib.check(lib.dfir_head_conv(x.data_ptr(), self.head_w[j].data_ptr(), self.head_b[j].data_ptr(),
                                          Hs[j].data_ptr(), None, B, self.in_feats, H, W, 64, st), "wide head")
            _lib.check(lib.dfir_stream_encode_hl8(Hs[j].data_ptr(), Hhi[j].data_ptr(), Hlo[j].data_ptr(), npl, st), "encode")
            w1, b1, w2, b2 = self.meta[j]
            _lib.check(lib.dfir_meta_attention(attr.data_ptr(), w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(),
                                               bufs["sq"][j].data_ptr(), nblk, B, self.M, self.hid, 64, self.meta_relu,
                                               None, self.res_scale, st), "wide meta")
        Pp, Zp = (Phi, Plo), (Zhi, Zlo)
        for k, ((w1t, b1), (w2t, b2)) in enumerate(self.blocks):
            xin = Hhi if k == 0 else Xhi
            for j in range(nc):  # t_j = relu(sum_i conv(x_i, W1[j][i]) + b1_j)
                for i in range(nc):
                    last = i == nc - 1
                    acc8(xin[i], w1t[j][i], b1[j] if last else None, None, Zp if i == 0 else Pp,
                         (T[j], None) if last else Pp, 1 if last else 0)
            for j in range(nc):  # x_j <- sum_i conv(t_i, W2[j][i]) * s_j + b2_j * s_j + x_j  (in place on the stream planes)
                s_j = bufs["sq"][j][k]
                xj = (Xhi[j], Xlo[j])
                for i in range(nc):
                    acc8(T[i], w2t[j][i], b2[j] if i == nc - 1 else None, s_j,
                         ((Hhi[j], Hlo[j]) if k == 0 else xj) if i == 0 else xj, xj, 0)
        wft, bfin = self.final
        src = Xhi if nblk > 0 else Hhi
        for j in range(nc):  # trunk tail conv + head skip
            for i in range(nc):
                last = i == nc - 1
                acc8(src[i], wft[j][i], bfin[j] if last else None, None, (Hhi[j], Hlo[j]) if i == 0 else Pp,
                     (Fbf[j], None) if last else Pp, 0)
        cur, h, wd = Fbf, H, W
        for (wu, bu), U in zip(self.ups, bufs["U"]):
            oh, ow = h * r, wd * r
            for jc in range(nc):
                for s in range(r * r):
                    si, sj = divmod(s, r)
                    for i in range(nc - 1):
                        acc(cur[i], wu[jc][s][i], None, None, None if i == 0 else P, P, junk, 0, h, wd)
                    base = U[jc].data_ptr() + (si * ow + sj) * 128
                    _lib.check(lib.dfir_conv3x3_c64(cur[nc - 1].data_ptr(), 64, 0, wu[jc][s][nc - 1].data_ptr(),
                                                    bu[jc][s].data_ptr(), B, h, wd, 3, 64, base, r * 128, r * ow * 128,
                                                    oh * ow * 128, P.data_ptr(), None, None, 0, st), "wide upsampler")
            cur, h, wd = U, oh, ow
        out = torch.empty(B, self.out_feats, h, wd, device=dev, dtype=torch.float32)
        for i in range(nc):
            _lib.check(lib.dfir_conv3x3_c64_tail(cur[i].data_ptr(), self.tail_w[i].data_ptr(),
                                                 self.tail_b.data_ptr() if i == 0 else None, B, h, wd, self.out_feats,
                                                 out.data_ptr(), 0 if i == 0 else 1, st), "wide tail")
        return out

    def launch_count(self):
        nc, rr = self.nc, self.r * self.r
        return 3 * nc + len(self.blocks) * 2 * nc * nc + nc * nc + len(self.ups) * nc * rr * nc + nc
