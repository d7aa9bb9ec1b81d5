// Channel-attention vector of QCALayer (all styles) as a cooperative device routine, usable from any
// thread group: `G` supplies the group's thread id, size and barrier.
//   reference: /root/reference/Code/SISR/models/attention_manipulators/architectures.py:105-125
// parameter order inside `p` (fp32, contiguous) for C channels, R = C/reduction, M metadata entries:
//   standard/modulate : W1[R][C]     b1[R]   W2[C][R]       b2[C]
//   max_concat/softmax: W1[R][C+M]   b1[R]   W2[C][R]       b2[C]
//   mini_concat       : Wp[R][C]     bp[R]   W2[C][R+M]     b2[C]
//   extended_attention: W1[C/2][C+M] b1  W2[C/4][C/2+M] b2  W3[R][C/4+M] b3  W4[C][R] b4
#pragma once
#include "../../include/dfir.h"

namespace dfir {

struct BlockGroup {  // the whole thread block
  __device__ __forceinline__ int tid() const { return threadIdx.x; }
  __device__ __forceinline__ int size() const { return blockDim.x; }
  __device__ __forceinline__ void sync() const { __syncthreads(); }
};

struct NamedGroup {  // a subset of warps synchronised through a named barrier
  int t, n, id;
  __device__ __forceinline__ int tid() const { return t; }
  __device__ __forceinline__ int size() const { return n; }
  __device__ __forceinline__ void sync() const { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
};

// number of fp32 parameters of one block's QCALayer in the order above
__host__ __device__ inline int attn_param_count(int style, int C, int R, int M) {
  switch (style) {
    case DFIR_STYLE_STANDARD:
    case DFIR_STYLE_MODULATE: return R * C + R + C * R + C;
    case DFIR_STYLE_MAX_CONCAT:
    case DFIR_STYLE_SOFTMAX: return R * (C + M) + R + C * R + C;
    case DFIR_STYLE_MINI_CONCAT: return R * C + R + C * (R + M) + C;
    case DFIR_STYLE_EXTENDED:
      return (C / 2) * (C + M) + C / 2 + (C / 4) * (C / 2 + M) + C / 4 + R * (C / 4 + M) + R + C * R + C;
    default: return 0;
  }
}

// Every QCALayer style is a chain of fully connected layers on the pooled vector: layer l maps
// [a_l (nin) ; attributes (M) if cat] -> nout, ReLU on every output but the last layer's, whose sigmoid (and, per style,
// softmax over channels / multiplication by the attributes) gives the scale.  mini_concat applies its ReLU to the
// concatenation [pre_concat(y); attributes], i.e. also to the attributes (cat_relu).  Parameters are stored layer by
// layer as W[nout][nin (+M)], b[nout] — the order listed above.  Used by the backward kernels (train_kernels.cu).
struct AttnChain {
  int L;
  int nin[4], nout[4], cat[4], cat_relu[4];
  int woff[4], boff[4];  // offsets into the block's parameter blob
  int aoff[4], doff[4];  // offsets of a_l and of the pre-activation gradients d_l inside a per-image signal record
  int dzq_off, sig_size, n_params;
};

__host__ __device__ inline AttnChain make_attn_chain(int style, int C, int R, int M) {
  AttnChain ch{};
  auto layer = [&](int l, int nin, int nout, int cat, int cat_relu) {
    ch.nin[l] = nin; ch.nout[l] = nout; ch.cat[l] = cat; ch.cat_relu[l] = cat_relu;
  };
  switch (style) {
    case DFIR_STYLE_STANDARD:
    case DFIR_STYLE_MODULATE: ch.L = 2; layer(0, C, R, 0, 0); layer(1, R, C, 0, 0); break;
    case DFIR_STYLE_MAX_CONCAT:
    case DFIR_STYLE_SOFTMAX: ch.L = 2; layer(0, C, R, 1, 0); layer(1, R, C, 0, 0); break;
    case DFIR_STYLE_MINI_CONCAT: ch.L = 2; layer(0, C, R, 0, 0); layer(1, R, C, 1, 1); break;
    case DFIR_STYLE_EXTENDED:
      ch.L = 4; layer(0, C, C / 2, 1, 0); layer(1, C / 2, C / 4, 1, 0); layer(2, C / 4, R, 1, 0); layer(3, R, C, 0, 0); break;
    default: ch.L = 0; break;
  }
  int off = 0, so = 0;
  for (int l = 0; l < ch.L; ++l) {
    ch.woff[l] = off; off += ch.nout[l] * (ch.nin[l] + (ch.cat[l] ? M : 0));
    ch.boff[l] = off; off += ch.nout[l];
    ch.aoff[l] = so; so += ch.nin[l];
  }
  for (int l = 0; l < ch.L; ++l) { ch.doff[l] = so; so += ch.nout[l]; }
  ch.dzq_off = so;
  ch.sig_size = so + C;
  ch.n_params = off;
  return ch;
}

template <class G>
__device__ void fc_layer(const G& g, const float* __restrict__ w, const float* __restrict__ bias, const float* in_a,
                         int na, const float* in_b, int nb, float* out, int nout,
                         int act /*0 none, 1 relu, 2 sigmoid*/) {
  const int nin = na + nb;
  if (nin >= 32 && (g.size() & 31) == 0) {
    // Squeeze layers (many inputs, few outputs): one thread per output would leave most of the group idle behind a
    // chain of nin dependent loads.  Eight lanes share an output (inputs interleaved by 8, butterfly over the lanes);
    // the summation order depends on the layer shape only, so every kernel that evaluates the layer gets the same bits.
    const int sub = g.tid() & 7, per_pass = g.size() >> 3;
    for (int o0 = 0; o0 < nout; o0 += per_pass) {
      const int o = o0 + (g.tid() >> 3);
      float s = 0.f;
      if (o < nout) {
        const float* wr = w + static_cast<size_t>(o) * nin;
#pragma unroll 4
        for (int i = sub; i < na; i += 8) s = fmaf(wr[i], in_a[i], s);
#pragma unroll 4
        for (int i = sub; i < nb; i += 8) s = fmaf(wr[na + i], in_b[i], s);
      }
      s += __shfl_xor_sync(0xffffffffu, s, 4);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      if (o < nout && sub == 0) {
        s += bias[o];
        if (act == 1) s = fmaxf(s, 0.f);
        if (act == 2) s = 1.f / (1.f + expf(-s));
        out[o] = s;
      }
    }
  } else {
    for (int o = g.tid(); o < nout; o += g.size()) {
      const float* wr = w + static_cast<size_t>(o) * nin;
      float s = bias[o];
      for (int i = 0; i < na; ++i) s = fmaf(wr[i], in_a[i], s);
      for (int i = 0; i < nb; ++i) s = fmaf(wr[na + i], in_b[i], s);
      if (act == 1) s = fmaxf(s, 0.f);
      if (act == 2) s = 1.f / (1.f + expf(-s));
      out[o] = s;
    }
  }
  g.sync();
}

// y_s[C]: pooled means (in);  s_s[C]: attention scale (out);  attr_s[A]: this image's attributes;
// tmp: >= 64 + C + M floats of scratch.  All pointers are shared memory visible to the whole group.
template <class G>
__device__ void attn_vector(const G& g, int style, const float* __restrict__ p, int C, int R, int M,
                            const float* attr_s, const float* y_s, float* s_s, float* tmp) {
  if (style == DFIR_STYLE_STANDARD || style == DFIR_STYLE_MODULATE) {
    const float* W1 = p; const float* b1 = W1 + R * C; const float* W2 = b1 + R; const float* b2 = W2 + C * R;
    fc_layer(g, W1, b1, y_s, C, nullptr, 0, tmp, R, 1);
    fc_layer(g, W2, b2, tmp, R, nullptr, 0, s_s, C, 2);
    if (style == DFIR_STYLE_MODULATE) {
      for (int c = g.tid(); c < C; c += g.size()) s_s[c] *= attr_s[c];
      g.sync();
    }
  } else if (style == DFIR_STYLE_MAX_CONCAT || style == DFIR_STYLE_SOFTMAX) {
    const float* W1 = p; const float* b1 = W1 + R * (C + M); const float* W2 = b1 + R; const float* b2 = W2 + C * R;
    fc_layer(g, W1, b1, y_s, C, attr_s, M, tmp, R, 1);
    fc_layer(g, W2, b2, tmp, R, nullptr, 0, s_s, C, 2);
    if (style == DFIR_STYLE_SOFTMAX) {
      // softmax over the C channels applied AFTER the sigmoid (architectures.py:100-101,119-121)
      if (g.tid() == 0) {
        float mx = -1e30f;
        for (int c = 0; c < C; ++c) mx = fmaxf(mx, s_s[c]);
        float sum = 0.f;
        for (int c = 0; c < C; ++c) sum += expf(s_s[c] - mx);
        tmp[0] = mx;
        tmp[1] = sum;
      }
      g.sync();
      const float mx = tmp[0], sum = tmp[1];
      g.sync();
      for (int c = g.tid(); c < C; c += g.size()) s_s[c] = expf(s_s[c] - mx) / sum;
      g.sync();
    }
  } else if (style == DFIR_STYLE_MINI_CONCAT) {
    const float* Wp = p; const float* bp = Wp + R * C; const float* W2 = bp + R; const float* b2 = W2 + C * (R + M);
    fc_layer(g, Wp, bp, y_s, C, nullptr, 0, tmp, R, 0);
    // conv_du = Sequential(ReLU, Conv, Sigmoid) on cat(pre, attributes): the ReLU hits both parts
    for (int i = g.tid(); i < R + M; i += g.size()) {
      const float v = i < R ? tmp[i] : attr_s[i - R];
      tmp[64 + i] = fmaxf(v, 0.f);
    }
    g.sync();
    fc_layer(g, W2, b2, tmp + 64, R + M, nullptr, 0, s_s, C, 2);
  } else if (style == DFIR_STYLE_EXTENDED) {
    const int c2 = C / 2, c4 = C / 4;
    const float* W1 = p; const float* b1 = W1 + c2 * (C + M);
    const float* W2 = b1 + c2; const float* b2 = W2 + c4 * (c2 + M);
    const float* W3 = b2 + c4; const float* b3 = W3 + R * (c4 + M);
    const float* W4 = b3 + R; const float* b4 = W4 + C * R;
    float* t1 = tmp; float* t2 = tmp + c2; float* t3 = t2 + c4;
    fc_layer(g, W1, b1, y_s, C, attr_s, M, t1, c2, 1);
    fc_layer(g, W2, b2, t1, c2, attr_s, M, t2, c4, 1);
    fc_layer(g, W3, b3, t2, c4, attr_s, M, t3, R, 1);
    fc_layer(g, W4, b4, t3, R, nullptr, 0, s_s, C, 2);
  }
}

}  // namespace dfir
