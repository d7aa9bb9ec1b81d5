// CUDA-core kernels of the TRAINING step (BaseModel.run_train -> loss.backward(),
// /root/reference/Code/SISR/models/__init__.py:466-489): everything in the backward pass that is not a dense
// contraction.  These are HBM- or latency-bound: reductions of g*r for the channel-attention gradient, the
// attention-MLP backward on [B][C] vectors, elementwise gradient forming, the fp32 parity-mode weight gradient, the
// 3-channel head/tail weight gradients, and the batched (pointer-table driven) re-packing of the fp32 parameters
// into kernel layouts after every optimizer step.
#include "kernels.h"
#include "attn.cuh"
#include "launch.cuh"
#include "ptx.cuh"

#include <algorithm>
#include <cmath>

namespace dfir {

namespace {

inline int ok_or_cuda3() { return cudaGetLastError() == cudaSuccess ? DFIR_OK : DFIR_ERR_CUDA; }

template <typename T>
__device__ __forceinline__ void load8(const T* p, float (&v)[8]);
template <>
__device__ __forceinline__ void load8<float>(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  const float4 b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <>
__device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 raw = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(h[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}
template <typename T>
__device__ __forceinline__ void store8(T* p, const float (&v)[8]);
template <>
__device__ __forceinline__ void store8<float>(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
template <>
__device__ __forceinline__ void store8<__nv_bfloat16>(__nv_bfloat16* p, const float (&v)[8]) {
  __align__(16) __nv_bfloat162 pk[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) pk[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = *reinterpret_cast<uint4*>(pk);
}

// ------------------------------------------------------------------------------------------------
// parameter re-packing driven by device pointer tables (one launch per category instead of one per layer)
// ------------------------------------------------------------------------------------------------
// blockIdx.y = j-th packed tile.  Source tensor = tbl[j / per_src] (or `direct`), OIHW fp32 [cout][64][3][3].
//   forward  : tile row n <-> output channel co = (j % per_src) + n * per_src, column k <-> input channel
//   transpose: the data-gradient conv: row n <-> INPUT channel, column k <-> output channel of the slice, taps
//              mirrored (correlation with the 180-degree rotated kernel)
__global__ void pack_bf16_multi_kernel(const float* const* __restrict__ tbl, const float* __restrict__ direct,
                                       __nv_bfloat16* __restrict__ out, int cout, int nt_rows, int per_src,
                                       int transpose, int j0) {
  const int j = blockIdx.y + j0;
  const float* w = tbl != nullptr ? tbl[j / per_src] : direct;
  const int slice = j % per_src;
  __nv_bfloat16* o = out + static_cast<size_t>(blockIdx.y) * 9 * nt_rows * 64;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 9 * nt_rows * 64) return;
  const int k = idx % 64;
  const int n = (idx / 64) % nt_rows;
  const int t = idx / (64 * nt_rows);
  float v = 0.f;
  if (!transpose) {
    const int co = slice + n * per_src;
    if (co < cout) v = w[(static_cast<size_t>(co) * 64 + k) * 9 + t];
  } else {
    const int co = slice + k * per_src;
    if (co < cout && n < 64) v = w[(static_cast<size_t>(co) * 64 + n) * 9 + (8 - t)];
  }
  const int chunk = (k >> 3) ^ (n & 7);
  o[static_cast<size_t>(t) * nt_rows * 64 + n * 64 + chunk * 8 + (k & 7)] = __float2bfloat16_rn(v);
}

// fp32 layouts: forward [9][Cin][Cout]; transpose [9][Cout][Cin] with mirrored taps (the dgrad conv's weights)
__global__ void pack_f32_multi_kernel(const float* const* __restrict__ tbl, const float* __restrict__ direct,
                                      float* __restrict__ out, int cout, int cin, int transpose) {
  const int j = blockIdx.y;
  const float* w = tbl != nullptr ? tbl[j] : direct;
  float* o = out + static_cast<size_t>(j) * 9 * cin * cout;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= 9 * cin * cout) return;
  if (!transpose) {
    const int co = idx % cout, ci = (idx / cout) % cin, t = idx / (cout * cin);
    o[idx] = w[(static_cast<size_t>(co) * cin + ci) * 9 + t];
  } else {
    const int ci = idx % cin, co = (idx / cin) % cout, t = idx / (cout * cin);
    o[idx] = w[(static_cast<size_t>(co) * cin + ci) * 9 + (8 - t)];
  }
}

// out[j*out_stride + i] = src[s + i*src_stride], src = tbl[(j / per_src)*tbl_stride + tbl_off] (NULL entry -> 0),
// s = j % per_src
__global__ void gather_strided_kernel(const float* const* __restrict__ tbl, const float* __restrict__ direct,
                                      int tbl_stride, int tbl_off, float* __restrict__ out, int n, int per_src,
                                      int src_stride, long long out_stride) {
  const int j = blockIdx.y;
  const float* src = tbl != nullptr ? tbl[static_cast<size_t>(j / per_src) * tbl_stride + tbl_off] : direct;
  const int s = j % per_src;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    out[static_cast<size_t>(j) * out_stride + i] = src != nullptr ? src[s + static_cast<size_t>(i) * src_stride] : 0.f;
}

// ------------------------------------------------------------------------------------------------
// part[b][chunk][c] = sum over the chunk's pixels of g * r      (d loss / d attention scale, before the batch of
// tiny MLP backward kernels).  HBM-bound: reads 4 + sizeof(T) bytes per element once, fully coalesced.
// ------------------------------------------------------------------------------------------------
struct CaBwdArgs;
__device__ void ca_backward_image(const CaBwdArgs& a, const int b, const int tid, const int NT);

template <typename T>
__device__ __forceinline__ void bwd_reduce_gr_body(const float* __restrict__ g, const T* __restrict__ r,
                                                   float* __restrict__ part, int HW, int C, int nchunk) {
  __shared__ float sh[2048];
  const int b = blockIdx.y, chunk = blockIdx.x, tid = threadIdx.x;
  const int lpp = C / 8, npl = 256 / lpp;
  const int c0 = (tid % lpp) * 8, pl = tid / lpp;
  const int ppc = (HW + nchunk - 1) / nchunk;
  const int p0 = chunk * ppc, p1 = min(HW, p0 + ppc);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const size_t img = static_cast<size_t>(b) * HW * C;
  for (int p = p0 + pl; p < p1; p += npl) {
    const size_t e = img + static_cast<size_t>(p) * C + c0;
    float gv[8], rv[8];
    load8<float>(g + e, gv);
    load8<T>(r + e, rv);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = fmaf(gv[i], rv[i], acc[i]);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) sh[pl * C + c0 + i] = acc[i];
  __syncthreads();
  for (int c = tid; c < C; c += 256) {
    float t = 0.f;
    for (int k = 0; k < npl; ++k) t += sh[k * C + c];
    part[(static_cast<size_t>(b) * nchunk + chunk) * C + c] = t;
  }
}

// ------------------------------------------------------------------------------------------------
// Backward of the block's scale vector s = QCALayer(mean(r), attributes) * meta_scale for image b = blockIdx.x
// (attention_manipulators/architectures.py:105-127, q_layer.py:39-43), every style (the MLP is walked as the generic
// layer chain of attn.cuh).  Emits what the two full-tensor passes need (s and the constant dL/dr contribution of the
// mean) and the per-image signals from which attn_param_grads_kernel forms the parameter gradients in a fixed order.
//   sig[b] = { a_0 = y, a_1, .., a_{L-1} | d_0, .., d_{L-1} | dzq[C] }     (AttnChain::aoff / doff / dzq_off)
// ------------------------------------------------------------------------------------------------
struct CaBwdArgs {
  const float* part; int nchunk;
  const float* pool_rows; int pool_nrows; int HW;
  int style, C, R, M, A;
  const float* ca;          // W1[R][Cin] b1[R] W2[C][R] b2[C]
  const float* attributes;  // [B][A]
  const float* sq;          // [B][C] meta scale of this block incl. out_scale, or nullptr
  float out_scale;          // factor folded into sq (ParamResBlock res_scale), 1 otherwise
  float* svec; float* dyv; float* sig; int sig_stride;
  const float* ymean;       // [B][C] pooled means saved by the forward (optional; else rebuilt from pool_rows)
};

__device__ void ca_backward_image(const CaBwdArgs& a, const int b, const int tid, const int NT) {
  __shared__ float ds[256], y[256], zl[256], sca[256], attr[512], tmp[512], dv[1024], red2[4];
  const int C = a.C, R = a.R;
  for (int c = tid; c < C; c += NT) {
    float t = 0.f;
    for (int k = 0; k < a.nchunk; ++k) t += __ldcg(&a.part[(static_cast<size_t>(b) * a.nchunk + k) * C + c]);
    ds[c] = t;
  }
  float* sig = a.sig + static_cast<size_t>(b) * a.sig_stride;
  if (a.style == DFIR_STYLE_NONE) {
    __syncthreads();
    for (int c = tid; c < C; c += NT) {
      const float sqv = a.sq != nullptr ? a.sq[static_cast<size_t>(b) * C + c] : a.out_scale;
      a.svec[static_cast<size_t>(b) * C + c] = sqv;
      const float sg = sqv / a.out_scale;  // sigmoid output
      sig[c] = a.sq != nullptr ? ds[c] * a.out_scale * sg * (1.f - sg) : 0.f;  // style none: the record is dzq only
    }
    return;
  }
  const AttnChain ch = make_attn_chain(a.style, C, R, a.M);
  const int M = a.M;
  float* av = tmp;   // activations a_0 .. a_{L-1}, laid out like the signal record (<= C + C/2 + C/4 + R floats)
  // pooled mean: saved by the forward, or rebuilt in the same fixed summation order as the forward streamer
  if (a.ymean != nullptr) {
    for (int c = tid; c < C; c += NT) y[c] = a.ymean[static_cast<size_t>(b) * C + c];
    for (int i = tid; i < a.A; i += NT) attr[i] = a.attributes[static_cast<size_t>(b) * a.A + i];
    __syncthreads();
  } else {
    const int ngrp = 256 / C;
    for (int i = tid; i < ngrp * C; i += NT) {
      const int c = i % C, grp = i / C;
      const float* pr = a.pool_rows + static_cast<size_t>(b) * a.pool_nrows * C + c;
      float s = 0.f;
      for (int row = grp; row < a.pool_nrows; row += ngrp) s += pr[static_cast<size_t>(row) * C];
      dv[i] = s;
    }
    for (int i = tid; i < a.A; i += NT) attr[i] = a.attributes[static_cast<size_t>(b) * a.A + i];
    __syncthreads();
    for (int c = tid; c < C; c += NT) {
      float t = 0.f;
      for (int gI = 0; gI < ngrp; ++gI) t += dv[gI * C + c];
      y[c] = t / static_cast<float>(a.HW);
    }
    __syncthreads();
  }
  for (int c = tid; c < C; c += NT) av[ch.aoff[0] + c] = y[c];
  __syncthreads();
  // ---- forward through the chain (hidden activations are needed by the backward)
  for (int l = 0; l < ch.L; ++l) {
    const int nin = ch.nin[l], kin = nin + (ch.cat[l] ? M : 0);
    const float* W = a.ca + ch.woff[l];
    const float* bias = a.ca + ch.boff[l];
    const float* in = av + ch.aoff[l];
    float* out = (l + 1 < ch.L) ? av + ch.aoff[l + 1] : zl;
    for (int o = tid; o < ch.nout[l]; o += NT) {
      const float* wr = W + static_cast<size_t>(o) * kin;
      float s = bias[o];
      for (int i = 0; i < nin; ++i) s = fmaf(wr[i], in[i], s);
      if (ch.cat[l])
        for (int m = 0; m < M; ++m) s = fmaf(wr[nin + m], ch.cat_relu[l] ? fmaxf(attr[m], 0.f) : attr[m], s);
      out[o] = (l + 1 < ch.L) ? fmaxf(s, 0.f) : s;
    }
    __syncthreads();
  }
  // ---- scale of the block and the gradient arriving at the last layer's pre-activation
  for (int c = tid; c < C; c += NT) zl[c] = 1.f / (1.f + expf(-zl[c]));  // sigmoid
  __syncthreads();
  if (a.style == DFIR_STYLE_SOFTMAX) {  // softmax over the channels AFTER the sigmoid (architectures.py:100-101)
    if (tid == 0) {
      float mx = -1e30f;
      for (int c = 0; c < C; ++c) mx = fmaxf(mx, zl[c]);
      float sum = 0.f;
      for (int c = 0; c < C; ++c) sum += expf(zl[c] - mx);
      red2[0] = mx;
      red2[1] = sum;
    }
    __syncthreads();
  }
  float* dl = dv + ch.doff[ch.L - 1] - ch.doff[0];  // d_l stored contiguously from dv[0] in signal order
  for (int c = tid; c < C; c += NT) {
    const float sg = zl[c];
    const float sqv = a.sq != nullptr ? a.sq[static_cast<size_t>(b) * C + c] : 1.f;
    float s_ca = sg;
    if (a.style == DFIR_STYLE_MODULATE) s_ca = sg * attr[c];
    if (a.style == DFIR_STYLE_SOFTMAX) s_ca = expf(sg - red2[0]) / red2[1];
    sca[c] = s_ca;
    a.svec[static_cast<size_t>(b) * C + c] = s_ca * sqv;
    const float sgq = sqv / a.out_scale;
    sig[ch.dzq_off + c] = a.sq != nullptr ? ds[c] * s_ca * a.out_scale * sgq * (1.f - sgq) : 0.f;
    ds[c] = ds[c] * sqv;  // now dL/d s_ca
  }
  __syncthreads();
  if (a.style == DFIR_STYLE_SOFTMAX) {
    if (tid == 0) {
      float dot = 0.f;
      for (int c = 0; c < C; ++c) dot = fmaf(sca[c], ds[c], dot);
      red2[2] = dot;
    }
    __syncthreads();
  }
  for (int c = tid; c < C; c += NT) {
    const float sg = zl[c];
    float d_sg = ds[c];
    if (a.style == DFIR_STYLE_MODULATE) d_sg *= attr[c];
    if (a.style == DFIR_STYLE_SOFTMAX) d_sg = sca[c] * (ds[c] - red2[2]);
    dl[c] = d_sg * sg * (1.f - sg);
  }
  __syncthreads();
  // ---- backward through the chain: d_{l-1} = (W_l[:, :nin]^T d_l) * (a_l > 0);  dy = W_0[:, :C]^T d_0
  for (int l = ch.L - 1; l >= 0; --l) {
    const int nin = ch.nin[l], kin = nin + (ch.cat[l] ? M : 0), nout = ch.nout[l];
    const float* W = a.ca + ch.woff[l];
    const float* dcur = dv + ch.doff[l] - ch.doff[0];
    for (int i = tid; i < nin; i += NT) {
      float s = 0.f;
      for (int o = 0; o < nout; ++o) s = fmaf(W[static_cast<size_t>(o) * kin + i], dcur[o], s);
      if (l > 0) dv[ch.doff[l - 1] - ch.doff[0] + i] = av[ch.aoff[l] + i] > 0.f ? s : 0.f;
      else a.dyv[static_cast<size_t>(b) * C + i] = s / static_cast<float>(a.HW);
    }
    __syncthreads();
  }
  // ---- per-image signal record for attn_param_grads_kernel
  for (int i = tid; i < ch.doff[0]; i += NT) sig[i] = av[i];
  for (int i = tid; i < ch.dzq_off - ch.doff[0]; i += NT) sig[ch.doff[0] + i] = dv[i];
}

// One launch for "ds = sum g*r" and the attention-MLP backward: the chunk CTA that finishes an image last (atomic
// ticket per image, self-resetting) runs ca_backward_image for it.  Deterministic: the partials are combined in chunk
// order whoever does it.
template <typename T>
__global__ void __launch_bounds__(256)
bwd_reduce_ca_kernel(const float* __restrict__ g, const T* __restrict__ r, float* __restrict__ part, int HW, int C,
                     int nchunk, CaBwdArgs a, unsigned int* __restrict__ tickets) {
  // constants first (overlaps the previous kernel's tail under programmatic dependent launch)
  __shared__ float caw_s[1024];
  const int n_caw = attn_param_count(a.style, a.C, a.R, a.M);
  const bool caw_in_smem = n_caw > 0 && n_caw <= 1024;
  if (caw_in_smem)
    for (int i = threadIdx.x; i < n_caw; i += 256) caw_s[i] = a.ca[i];
  ptx::grid_dep_wait();  // (no early launch_dependents: a resident, waiting successor would starve our later waves)
  bwd_reduce_gr_body<T>(g, r, part, HW, C, nchunk);
  __shared__ unsigned int last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(&tickets[blockIdx.y], 1u) == static_cast<unsigned>(nchunk - 1) ? 1u : 0u;
  __syncthreads();
  if (last == 0u) return;
  __threadfence();
  if (threadIdx.x == 0) tickets[blockIdx.y] = 0u;
  CaBwdArgs la = a;
  if (caw_in_smem) la.ca = caw_s;
  ca_backward_image(la, blockIdx.y, threadIdx.x, 256);
}

// dr = g * s[b][c] + dyv[b][c]   (dL/d conv2 output of the block), written in the operand format T
template <typename T>
__global__ void __launch_bounds__(256)
form_dr_kernel(const float* __restrict__ g, const float* __restrict__ svec, const float* __restrict__ dyv,
               T* __restrict__ dr, int HW, int C) {
  ptx::grid_dep_wait();  // (no early launch_dependents: a resident, waiting successor would starve our later waves)
  const int b = blockIdx.y, tid = threadIdx.x;
  const int lpp = C / 8;
  const int c0 = (tid % lpp) * 8;
  float sc[8], ad[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    sc[i] = svec[static_cast<size_t>(b) * C + c0 + i];
    ad[i] = dyv != nullptr ? dyv[static_cast<size_t>(b) * C + c0 + i] : 0.f;
  }
  const size_t img = static_cast<size_t>(b) * HW * C;
  const long long nvec = static_cast<long long>(HW) * lpp;
  for (long long vi = static_cast<long long>(blockIdx.x) * 256 + tid; vi < nvec;
       vi += static_cast<long long>(gridDim.x) * 256) {
    const size_t e = img + static_cast<size_t>(vi) * 8;
    float gv[8], o[8];
    load8<float>(g + e, gv);
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = fmaf(gv[i], sc[i], ad[i]);
    store8<T>(dr + e, o);
  }
}

// ------------------------------------------------------------------------------------------------
// Pixel attention backward (PALayer, attention_manipulators/architectures.py:13-26, inside QRCAB.forward :172-180):
//   u = r * s_ca,  pm = sigmoid(w2 . relu(W1 u + b1) + b2),  v = u * pm,  x' = v * sq + x        (C = 64, hidden 8)
// given g = dL/dx':  dv = g * sq;  A = sum_c dv_c u_c;  dz = A pm (1 - pm);  dh_j = dz w2_j [h_j > 0];
//   du_c = dv_c pm + sum_j W1[j][c] dh_j      -> written out: it is the `g` of the channel-attention backward that follows
//   dsq_c = sum_p g_c v_c;  dW1[j][c] = sum dh_j u_c;  db1 = sum dh;  dw2_j = sum dz relu(h_j);  db2 = sum dz
// Each CTA writes one record of partial sums { dW1[8][64], db1[8], dw2[8], db2, dsq[64] }; pa_finish_kernel combines them in
// a fixed order.  The 8 threads of a pixel hold 8 channels each (butterfly over the lanes, as in the forward streamer).
// ------------------------------------------------------------------------------------------------
constexpr int kPaParams = 8 * 64 + 8 + 8 + 1;
constexpr int kPaRec = kPaParams + 64;

template <typename T>
__global__ void __launch_bounds__(256)
pa_backward_kernel(const float* __restrict__ g, const T* __restrict__ r, const float* __restrict__ ymean, AttnParams ap,
                   const float* __restrict__ attributes, const float* __restrict__ sq, const float* __restrict__ pa,
                   float* __restrict__ du, float* __restrict__ part, int HW) {
  constexpr int C = 64;
  __shared__ float y_s[256], s_s[256], attr_s[512], tmp[4 * 256];
  __shared__ float pa_s[kPaParams + 3];
  __shared__ float red[32 * 64];
  const int b = blockIdx.y, tid = threadIdx.x;
  for (int i = tid; i < kPaParams; i += 256) pa_s[i] = pa[i];
  for (int i = tid; i < ap.A; i += 256) attr_s[i] = attributes[static_cast<size_t>(b) * ap.A + i];
  if (tid < C) y_s[tid] = ymean[static_cast<size_t>(b) * C + tid];
  __syncthreads();
  attn_vector(BlockGroup{}, ap.style, ap.w[0], C, ap.R, ap.M, attr_s, y_s, s_s, tmp);
  __syncthreads();
  const int lane8 = tid & 7, pg = tid >> 3;
  const int c0 = lane8 * 8;
  float sc[8], sqv[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    sc[i] = s_s[c0 + i];
    sqv[i] = sq != nullptr ? sq[static_cast<size_t>(b) * C + c0 + i] : 1.f;
  }
  float aW1[8][8], ab1[8], aw2[8], ab2 = 0.f, adsq[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    ab1[j] = 0.f; aw2[j] = 0.f; adsq[j] = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) aW1[j][i] = 0.f;
  }
  const size_t img_off = static_cast<size_t>(b) * HW * C;
  const long long nvec = static_cast<long long>(HW) * 8;
  for (long long base = static_cast<long long>(blockIdx.x) * 256; base < nvec; base += static_cast<long long>(gridDim.x) * 256) {
    const long long vi = base + tid;
    const bool active = vi < nvec;
    const size_t e = img_off + static_cast<size_t>(active ? vi : 0) * 8;
    float u[8], gv[8];
    load8<T>(r + e, u);
    load8<float>(g + e, gv);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      u[i] *= sc[i];
      if (!active) gv[i] = 0.f;
    }
    float h[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float t = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) t = fmaf(pa_s[j * 64 + c0 + i], u[i], t);
      h[j] = t;
    }
#pragma unroll
    for (int mask = 1; mask < 8; mask <<= 1)
#pragma unroll
      for (int j = 0; j < 8; ++j) h[j] += __shfl_xor_sync(0xffffffffu, h[j], mask);
    float z = pa_s[8 * 64 + 16];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      h[j] += pa_s[8 * 64 + j];
      z = fmaf(pa_s[8 * 64 + 8 + j], fmaxf(h[j], 0.f), z);
    }
    const float pm = 1.f / (1.f + expf(-z));
    float A = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) A = fmaf(gv[i] * sqv[i], u[i], A);
#pragma unroll
    for (int mask = 1; mask < 8; mask <<= 1) A += __shfl_xor_sync(0xffffffffu, A, mask);
    const float dz = A * pm * (1.f - pm);
    float dh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) dh[j] = h[j] > 0.f ? dz * pa_s[8 * 64 + 8 + j] : 0.f;
    float o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float t = gv[i] * sqv[i] * pm;
#pragma unroll
      for (int j = 0; j < 8; ++j) t = fmaf(pa_s[j * 64 + c0 + i], dh[j], t);
      o[i] = t;
      adsq[i] = fmaf(gv[i] * u[i], pm, adsq[i]);
    }
    if (active) store8<float>(du + e, o);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
#pragma unroll
      for (int i = 0; i < 8; ++i) aW1[j][i] = fmaf(dh[j], u[i], aW1[j][i]);
      if (lane8 == 0) {
        ab1[j] += dh[j];
        aw2[j] = fmaf(dz, fmaxf(h[j], 0.f), aw2[j]);
      }
    }
    if (lane8 == 0) ab2 += dz;
  }
  // ---- CTA reduction over the 32 pixel groups, fixed order
  float* rec = part + (static_cast<size_t>(b) * gridDim.x + blockIdx.x) * kPaRec;
#pragma unroll
  for (int j = 0; j < 9; ++j) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i) red[pg * 64 + c0 + i] = j < 8 ? aW1[j < 8 ? j : 0][i] : adsq[i];
    __syncthreads();
    if (tid < 64) {
      float t = 0.f;
      for (int k = 0; k < 32; ++k) t += red[k * 64 + tid];
      rec[j < 8 ? j * 64 + tid : kPaParams + tid] = t;
    }
  }
  __syncthreads();
  if (lane8 == 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      red[pg * 17 + j] = ab1[j];
      red[pg * 17 + 8 + j] = aw2[j];
    }
    red[pg * 17 + 16] = ab2;
  }
  __syncthreads();
  if (tid < 17) {
    float t = 0.f;
    for (int k = 0; k < 32; ++k) t += red[k * 17 + tid];
    rec[8 * 64 + tid] = t;
  }
}

// parameter gradients (summed over images and CTAs in index order) and the meta-attention signal dzq of the block's record
__global__ void __launch_bounds__(128)
pa_finish_kernel(const float* __restrict__ part, int ncta, int B, const float* __restrict__ sq, float out_scale,
                 float* __restrict__ sig, int sig_stride, int dzq_off, float* const* __restrict__ gpa) {
  const int idx = blockIdx.x * 128 + threadIdx.x;
  if (idx < kPaParams) {
    float t = 0.f;
    for (int k = 0; k < B * ncta; ++k) t += part[static_cast<size_t>(k) * kPaRec + idx];
    float* dst = idx < 512 ? gpa[0] + idx : (idx < 520 ? gpa[1] + (idx - 512) : (idx < 528 ? gpa[2] + (idx - 520) : gpa[3]));
    if (dst != nullptr) *dst = t;
  } else if (idx < kPaParams + 64) {
    const int c = idx - kPaParams;
    for (int b = 0; b < B; ++b) {
      float t = 0.f;
      for (int k = 0; k < ncta; ++k) t += part[(static_cast<size_t>(b) * ncta + k) * kPaRec + idx];
      const float sqv = sq != nullptr ? sq[static_cast<size_t>(b) * 64 + c] : out_scale;
      const float sgq = sqv / out_scale;
      sig[static_cast<size_t>(b) * sig_stride + dzq_off + c] = sq != nullptr ? t * out_scale * sgq * (1.f - sgq) : 0.f;
    }
  }
}

// out32 = a + b (either may alias out32), optional bf16 copy
__global__ void add_f32_kernel(const float4* __restrict__ a, const float4* __restrict__ b, float4* __restrict__ out,
                               uint2* __restrict__ out_bf16, long long n4) {
  ptx::grid_dep_wait();  // (no early launch_dependents: a resident, waiting successor would starve our later waves)
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 x = a[i], y = b[i];
    const float4 o = make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w);
    out[i] = o;
    if (out_bf16 != nullptr) {
      __nv_bfloat162 lo = __floats2bfloat162_rn(o.x, o.y), hi = __floats2bfloat162_rn(o.z, o.w);
      uint2 pk;
      pk.x = *reinterpret_cast<uint32_t*>(&lo);
      pk.y = *reinterpret_cast<uint32_t*>(&hi);
      out_bf16[i] = pk;
    }
  }
}

// inverse of nn.PixelShuffle(r) on NHWC fp32: in [B][h*r][w*r][C] -> out [B][h][w][C*r*r], channel c*r*r + i*r + j
__global__ void pixel_unshuffle_f32_kernel(const float* __restrict__ in, float* __restrict__ out, int B, int h, int w,
                                           int C, int r) {
  const long long n = static_cast<long long>(B) * h * w * C * r * r;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < n;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int k = static_cast<int>(idx % (C * r * r));
    const long long pix = idx / (C * r * r);
    const int x = static_cast<int>(pix % w), y = static_cast<int>((pix / w) % h), b = static_cast<int>(pix / (static_cast<long long>(w) * h));
    const int c = k / (r * r), ij = k % (r * r), i = ij / r, j = ij % r;
    out[idx] = in[((static_cast<size_t>(b) * h * r + (y * r + i)) * (static_cast<size_t>(w) * r) + (x * r + j)) * C + c];
  }
}

// ------------------------------------------------------------------------------------------------
// fp32 parity-mode weight gradient of a 3x3 conv: part[s][tap][ci][co] = sum over the CTA's rows of
// dY[p][co] * X[p + tap][ci] (zero padded), dbpart[s][co] = sum dY[p][co].
// grid (S row chunks, Cin/16, Cout/64), 256 threads: thread = (4 output channels, 1 input channel) x 9 taps.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
wgrad_f32_kernel(const float* __restrict__ dY, const float* __restrict__ X, float* __restrict__ part,
                 float* __restrict__ dbpart, int B, int H, int W, int Cin, int Cout) {
  __shared__ __align__(16) float dys[32][64];
  __shared__ float xs[3][34][16];
  const int tid = threadIdx.x;
  const int cq = tid & 15, ci = tid >> 4;
  const int S = gridDim.x, s = blockIdx.x;
  const int ci0 = blockIdx.y * 16, co0 = blockIdx.z * 64;
  const long long rows = static_cast<long long>(B) * H;
  const long long r0 = rows * s / S, r1 = rows * (s + 1) / S;
  float acc[9][4];
#pragma unroll
  for (int t = 0; t < 9; ++t)
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[t][k] = 0.f;
  float dbacc[4] = {0.f, 0.f, 0.f, 0.f};
  for (long long row = r0; row < r1; ++row) {
    const int b = static_cast<int>(row / H), y = static_cast<int>(row % H);
    for (int x0 = 0; x0 < W; x0 += 32) {
      for (int i = tid; i < 32 * 64; i += 256) {
        const int px = i >> 6, c = i & 63;
        const int x = x0 + px;
        dys[px][c] = (x < W && co0 + c < Cout) ? dY[((static_cast<size_t>(b) * H + y) * W + x) * Cout + co0 + c] : 0.f;
      }
      for (int i = tid; i < 3 * 34 * 16; i += 256) {
        const int c = i & 15, px = (i >> 4) % 34, dy = i / (34 * 16);
        const int yy = y + dy - 1, xx = x0 + px - 1;
        xs[dy][px][c] = (yy >= 0 && yy < H && xx >= 0 && xx < W && ci0 + c < Cin)
                            ? X[((static_cast<size_t>(b) * H + yy) * W + xx) * Cin + ci0 + c] : 0.f;
      }
      __syncthreads();
#pragma unroll 4
      for (int px = 0; px < 32; ++px) {
        const float4 d4 = *reinterpret_cast<const float4*>(&dys[px][cq * 4]);
        dbacc[0] += d4.x; dbacc[1] += d4.y; dbacc[2] += d4.z; dbacc[3] += d4.w;
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) {
            const float xv = xs[dy][px + dx][ci];
            acc[dy * 3 + dx][0] = fmaf(d4.x, xv, acc[dy * 3 + dx][0]);
            acc[dy * 3 + dx][1] = fmaf(d4.y, xv, acc[dy * 3 + dx][1]);
            acc[dy * 3 + dx][2] = fmaf(d4.z, xv, acc[dy * 3 + dx][2]);
            acc[dy * 3 + dx][3] = fmaf(d4.w, xv, acc[dy * 3 + dx][3]);
          }
      }
      __syncthreads();
    }
  }
  if (ci0 + ci < Cin) {
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int co = co0 + cq * 4 + k;
        if (co < Cout) part[((static_cast<size_t>(s) * 9 + t) * Cin + ci0 + ci) * Cout + co] = acc[t][k];
      }
  }
  if (blockIdx.y == 0 && ci == 0) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int co = co0 + cq * 4 + k;
      if (co < Cout) dbpart[static_cast<size_t>(s) * Cout + co] = dbacc[k];
    }
  }
}

// dW[co][ci][tap] (OIHW, co = co_begin + n*co_stride) = sum_s part[s][tap][ci][n];  db likewise.  Destinations are
// read from the gradient pointer tables on the device (w_tbl[w_idx] / b_tbl[b_idx]) or given directly.
// 256 threads = 64 outputs x 4 groups of partials; every group sums its partials s = grp, grp+4, ... with 8 loads in
// flight, the four group sums are combined in a fixed order (deterministic).  Output index space: the 9*Cin*n_rows
// weight entries followed by the n_rows bias entries.
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float* __restrict__ part, const float* __restrict__ dbpart, int S, int Cin, int n_rows,
                    float* const* __restrict__ w_tbl, int w_idx, float* w_direct, float* const* __restrict__ b_tbl,
                    int b_idx, float* b_direct, int co_begin, int co_stride, int co_major) {
  ptx::grid_dep_wait();  // (no early launch_dependents: a resident, waiting successor would starve our later waves)
  __shared__ float red[4][64];
  const int total = 9 * Cin * n_rows;
  const int o = threadIdx.x & 63, grp = threadIdx.x >> 6;
  const int idx = blockIdx.x * 64 + o;
  float t = 0.f;
  if (idx < total + n_rows) {
    const float* src = idx < total ? part + idx : dbpart + (idx - total);
    const size_t stride = idx < total ? static_cast<size_t>(total) : static_cast<size_t>(n_rows);
    int s = grp;
    for (; s + 28 < S; s += 32) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = src[static_cast<size_t>(s + 4 * u) * stride];
#pragma unroll
      for (int u = 0; u < 8; ++u) t += v[u];
    }
    for (; s < S; s += 4) t += src[static_cast<size_t>(s) * stride];
  }
  red[grp][o] = t;
  __syncthreads();
  if (grp != 0 || idx >= total + n_rows) return;
  const float sum = ((red[0][o] + red[1][o]) + red[2][o]) + red[3][o];
  if (idx < total) {
    float* dw = w_tbl != nullptr ? w_tbl[w_idx] : w_direct;
    // partial layout: [tap][ci][n] (CUDA-core / mma.sync kernels) or [tap][n][ci] (tcgen05 kernel)
    const int tap = idx / (n_rows * Cin);
    const int n = co_major ? (idx / Cin) % n_rows : idx % n_rows;
    const int ci = co_major ? idx % Cin : (idx / n_rows) % Cin;
    if (dw != nullptr) dw[(static_cast<size_t>(co_begin + n * co_stride) * Cin + ci) * 9 + tap] = sum;
  } else {
    float* db = b_tbl != nullptr ? b_tbl[b_idx] : b_direct;
    if (db != nullptr) db[co_begin + (idx - total) * co_stride] = sum;
  }
}

// ------------------------------------------------------------------------------------------------
// weight gradients of the two 3-channel convs (head 3 -> C, tail C -> 3): correlation of a small NCHW fp32 image
// I[B][3][H][W] with an NHWC feature map F[B][H][W][C]:
//     K[c3][t][c] = sum_q F[q][c] * I[c3][q + off(t)],   Fsum[c] = sum_q F[q][c],   Isum[c3] = sum_q I[c3][q]
// grid (S row chunks, C/64), 256 threads = 64 channels x 4 pixel lanes.  Partials are reduced by
// wgrad_small_reduce_kernel:  head  dW[co][ci][tap] = K[ci][tap][co],   db = Fsum
//                             tail  dW[co][ci][tap] = K[co][8-tap][ci], db = Isum
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
wgrad_small_kernel(const float* __restrict__ I, const T* __restrict__ F, float* __restrict__ kpart,
                   float* __restrict__ fsum_part, float* __restrict__ isum_part, int B, int H, int W, int C, int C3) {
  extern __shared__ float sm[];  // is[C3][3][W+2] then red[4][28][64]
  float* is = sm;
  float* red = sm + 3 * 3 * (W + 2);
  const int tid = threadIdx.x, c = tid & 63, pl = tid >> 6;
  const int S = gridDim.x, s = blockIdx.x, cb = blockIdx.y * 64;
  const long long rows = static_cast<long long>(B) * H;
  const long long r0 = rows * s / S, r1 = rows * (s + 1) / S;
  float acc[27];
#pragma unroll
  for (int i = 0; i < 27; ++i) acc[i] = 0.f;
  float fs = 0.f, isum = 0.f;
  const int Wp = W + 2;
  for (long long row = r0; row < r1; ++row) {
    const int b = static_cast<int>(row / H), y = static_cast<int>(row % H);
    __syncthreads();
    for (int i = tid; i < 9 * Wp; i += 256) {
      const int px = i % Wp, dy = (i / Wp) % 3, c3 = i / (3 * Wp);
      const int yy = y + dy - 1, xx = px - 1;
      is[i] = (c3 < C3 && yy >= 0 && yy < H && xx >= 0 && xx < W) ? I[((static_cast<size_t>(b) * C3 + c3) * H + yy) * W + xx] : 0.f;
    }
    __syncthreads();
    if (tid < C3) {
      float t = 0.f;
      for (int x = 0; x < W; ++x) t += is[(tid * 3 + 1) * Wp + x + 1];
      isum += t;
    }
    const T* frow = F + (static_cast<size_t>(b) * H + y) * W * C + cb + c;
    for (int x = pl; x < W; x += 4) {
      const float f = static_cast<float>(frow[static_cast<size_t>(x) * C]);
      fs += f;
#pragma unroll
      for (int c3 = 0; c3 < 3; ++c3)
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) acc[c3 * 9 + dy * 3 + dx] = fmaf(f, is[(c3 * 3 + dy) * Wp + x + dx], acc[c3 * 9 + dy * 3 + dx]);
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 27; ++i) red[(pl * 28 + i) * 64 + c] = acc[i];
  red[(pl * 28 + 27) * 64 + c] = fs;
  __syncthreads();
  if (pl == 0) {
    for (int i = 0; i < 28; ++i) {
      const float t = ((red[(0 * 28 + i) * 64 + c] + red[(1 * 28 + i) * 64 + c]) + red[(2 * 28 + i) * 64 + c]) +
                      red[(3 * 28 + i) * 64 + c];
      if (i < 27) kpart[(static_cast<size_t>(s) * 27 + i) * C + cb + c] = t;
      else fsum_part[static_cast<size_t>(s) * C + cb + c] = t;
    }
  }
  if (tid < C3 && blockIdx.y == 0) isum_part[static_cast<size_t>(s) * 4 + tid] = isum;
}

__global__ void wgrad_small_reduce_kernel(const float* __restrict__ kpart, const float* __restrict__ fsum_part,
                                          const float* __restrict__ isum_part, int S, int C, int C3, int tail_mode,
                                          float* __restrict__ dw, float* __restrict__ db) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int total = C3 * 9 * C;
  if (idx < total) {
    const int c = idx % C, t = (idx / C) % 9, c3 = idx / (9 * C);
    float v = 0.f;
    for (int s = 0; s < S; ++s) v += kpart[(static_cast<size_t>(s) * 27 + c3 * 9 + t) * C + c];
    if (tail_mode) dw[(static_cast<size_t>(c3) * C + c) * 9 + (8 - t)] = v;   // [co=c3][ci=c][tap]
    else dw[(static_cast<size_t>(c) * C3 + c3) * 9 + t] = v;                   // [co=c][ci=c3][tap]
  }
  const int nb = tail_mode ? C3 : C;
  if (idx < nb) {
    float v = 0.f;
    for (int s = 0; s < S; ++s) v += tail_mode ? isum_part[static_cast<size_t>(s) * 4 + idx] : fsum_part[static_cast<size_t>(s) * C + idx];
    db[idx] = v;
  }
}

// ------------------------------------------------------------------------------------------------
// Parameter gradients of every block's attention MLPs from the stored per-image signals: grid = nblk CTAs,
// each loops over the batch in a fixed order (deterministic).  Dynamic smem: hq[B][Hid], dhq[B][Hid].
// ------------------------------------------------------------------------------------------------
struct AttnGradArgs {
  const float* sig; int sig_stride;      // [nblk][B][sig_stride]
  const float* attributes; int A;        // [B][A]
  const float* meta_w1; const float* meta_b1; const float* meta_w2;  // packed [nblk][Hid][M], [nblk][Hid], [nblk][C][Hid]
  const int* q_enabled;
  float* const* ca_g;                    // [nblk*8] (W, b) gradients per chain layer (nullptr table = no channel attention)
  float* const* meta_g;                  // [nblk*4] FC1 w, b, FC2 w, b gradients (NULL entries where no q layer)
  int B, C, R, M, Hid, style, meta_relu;
};

__global__ void __launch_bounds__(256) attn_param_grads_kernel(AttnGradArgs a) {
  extern __shared__ float sm[];
  const int blk = blockIdx.x, tid = threadIdx.x;
  const int B = a.B, C = a.C, R = a.R, M = a.M, Hid = a.Hid;
  const float* sig = a.sig + static_cast<size_t>(blk) * B * a.sig_stride;
  const AttnChain ch = make_attn_chain(a.style, C, R, M);
  const int o_dzq = ch.dzq_off;
  if (a.style != DFIR_STYLE_NONE && a.ca_g != nullptr) {
    for (int l = 0; l < ch.L; ++l) {  // dW_l[o][i] = sum_b d_l[b][o] * in_l[b][i],  db_l[o] = sum_b d_l[b][o]
      float* gW = a.ca_g[blk * 8 + 2 * l];
      float* gb = a.ca_g[blk * 8 + 2 * l + 1];
      const int nin = ch.nin[l], kin = nin + (ch.cat[l] ? M : 0), nout = ch.nout[l];
      for (int e = tid; e < nout * kin; e += 256) {
        const int o = e / kin, i = e % kin;
        float t = 0.f;
        for (int b = 0; b < B; ++b) {
          float in;
          if (i < nin) in = sig[b * a.sig_stride + ch.aoff[l] + i];
          else {
            in = a.attributes[static_cast<size_t>(b) * a.A + (i - nin)];
            if (ch.cat_relu[l]) in = fmaxf(in, 0.f);
          }
          t = fmaf(sig[b * a.sig_stride + ch.doff[l] + o], in, t);
        }
        gW[e] = t;
      }
      for (int o = tid; o < nout; o += 256) {
        float t = 0.f;
        for (int b = 0; b < B; ++b) t += sig[b * a.sig_stride + ch.doff[l] + o];
        gb[o] = t;
      }
    }
  }
  if (a.meta_g == nullptr || a.q_enabled == nullptr || a.q_enabled[blk] == 0) return;
  float* gV1 = a.meta_g[blk * 4 + 0]; float* gc1 = a.meta_g[blk * 4 + 1];
  float* gV2 = a.meta_g[blk * 4 + 2]; float* gc2 = a.meta_g[blk * 4 + 3];
  if (gV1 == nullptr) return;
  const float* V1 = a.meta_w1 + static_cast<size_t>(blk) * Hid * M;
  const float* c1 = a.meta_b1 + static_cast<size_t>(blk) * Hid;
  const float* V2 = a.meta_w2 + static_cast<size_t>(blk) * C * Hid;
  float* hq = sm;             // [B][Hid]
  float* dhq = sm + B * Hid;  // [B][Hid]
  for (int e = tid; e < B * Hid; e += 256) {
    const int b = e / Hid, j = e % Hid;
    float s = c1[j];
    for (int i = 0; i < M; ++i) s = fmaf(V1[j * M + i], a.attributes[static_cast<size_t>(b) * a.A + i], s);
    const float hv = a.meta_relu ? fmaxf(s, 0.f) : s;
    hq[e] = hv;
    float d = 0.f;
    for (int c = 0; c < C; ++c) d = fmaf(V2[static_cast<size_t>(c) * Hid + j], sig[b * a.sig_stride + o_dzq + c], d);
    dhq[e] = (a.meta_relu && !(s > 0.f)) ? 0.f : d;
  }
  __syncthreads();
  for (int e = tid; e < C * Hid; e += 256) {
    const int c = e / Hid, j = e % Hid;
    float t = 0.f;
    for (int b = 0; b < B; ++b) t = fmaf(sig[b * a.sig_stride + o_dzq + c], hq[b * Hid + j], t);
    gV2[e] = t;
  }
  for (int c = tid; c < C; c += 256) {
    float t = 0.f;
    for (int b = 0; b < B; ++b) t += sig[b * a.sig_stride + o_dzq + c];
    gc2[c] = t;
  }
  for (int e = tid; e < Hid * M; e += 256) {
    const int j = e / M, i = e % M;
    float t = 0.f;
    for (int b = 0; b < B; ++b) t = fmaf(dhq[b * Hid + j], a.attributes[static_cast<size_t>(b) * a.A + i], t);
    gV1[e] = t;
  }
  for (int j = tid; j < Hid; j += 256) {
    float t = 0.f;
    for (int b = 0; b < B; ++b) t += dhq[b * Hid + j];
    gc1[j] = t;
  }
}

// ------------------------------------------------------------------------------------------------
// Adam over flat fp32 buffers (torch.optim.Adam semantics, `BaseModel.standard_update`,
// /root/reference/Code/SISR/models/__init__.py:481-489): one pass over parameters, gradients and both moments
//   m = b1 m + (1-b1) g ;  v = b2 v + (1-b2) g^2 ;  p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
// HBM-bound: 16 B read + 12 B written per parameter.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
adam_flat_kernel(float4* __restrict__ p, const float4* __restrict__ g, float4* __restrict__ m, float4* __restrict__ v,
                 long long n4, float lr, float b1, float b2, float eps, float wd, float bc1, float bc2_sqrt) {
  const float step_size = lr / bc1;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float4 pp = p[i], gg = g[i], mm = m[i], vv = v[i];
    float* pa = &pp.x; float* ga = &gg.x; float* ma = &mm.x; float* va = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float grad = ga[k] + wd * pa[k];
      ma[k] = ma[k] + (1.f - b1) * (grad - ma[k]);              // lerp form, as torch's fused kernel
      va[k] = b2 * va[k] + (1.f - b2) * grad * grad;
      const float denom = sqrtf(va[k]) / bc2_sqrt + eps;
      pa[k] = pa[k] - step_size * (ma[k] / denom);
    }
    p[i] = pp; m[i] = mm; v[i] = vv;
  }
}

unsigned grid_for(long long n, int per_block = 256, long long cap = 148 * 16) {
  long long g = (n + per_block - 1) / per_block;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<unsigned>(g);
}

}  // namespace

// ------------------------------------------------------------------------------------------------ host wrappers
int pack_bf16_multi(const float* const* tbl, const float* direct, void* out, int n_tiles, int cout, int nt_rows,
                    int per_src, int transpose, cudaStream_t s, int j0) {
  if (n_tiles <= 0) return DFIR_OK;
  dim3 grid((9 * nt_rows * 64 + 255) / 256, n_tiles);
  pack_bf16_multi_kernel<<<grid, 256, 0, s>>>(tbl, direct, reinterpret_cast<__nv_bfloat16*>(out), cout, nt_rows, per_src,
                                              transpose, j0);
  return ok_or_cuda3();
}

int pack_f32_multi(const float* const* tbl, const float* direct, float* out, int n, int cout, int cin, int transpose,
                   cudaStream_t s) {
  if (n <= 0) return DFIR_OK;
  dim3 grid((9 * cin * cout + 255) / 256, n);
  pack_f32_multi_kernel<<<grid, 256, 0, s>>>(tbl, direct, out, cout, cin, transpose);
  return ok_or_cuda3();
}

int gather_strided(const float* const* tbl, const float* direct, int tbl_stride, int tbl_off, float* out, int n_rows,
                   int n, int per_src, int src_stride, long long out_stride, cudaStream_t s) {
  if (n_rows <= 0 || n <= 0) return DFIR_OK;
  dim3 grid(std::min((n + 255) / 256, 64), n_rows);
  gather_strided_kernel<<<grid, 256, 0, s>>>(tbl, direct, tbl_stride, tbl_off, out, n, per_src, src_stride, out_stride);
  return ok_or_cuda3();
}

int bwd_reduce_chunks(int HW) { return std::max(1, std::min(32, HW / 256)); }

int bwd_reduce_ca(const float* g, const void* r, int r_is_bf16, float* part, unsigned int* tickets,
                  const float* pool_rows, int pool_nrows, const float* ymean, int HW, const AttnParams& ap,
                  const float* attributes, const float* sq, float out_scale, float* svec, float* dyv, float* sig,
                  int sig_stride, int B, cudaStream_t s) {
  const int C = ap.C;
  if (C % 8 != 0 || C > 256 || 256 % (C / 8) != 0) return DFIR_ERR_ARG;
  if (ap.C > 256 || 256 % ap.C != 0 || ap.R > 64 || ap.A > 512) return DFIR_ERR_ARG;
  if (ap.style < DFIR_STYLE_NONE || ap.style > DFIR_STYLE_EXTENDED) return DFIR_ERR_ARG;
  if (sig_stride < make_attn_chain(ap.style, ap.C, ap.R, ap.M).sig_size) return DFIR_ERR_ARG;
  CaBwdArgs a{};
  a.part = part; a.nchunk = bwd_reduce_chunks(HW); a.pool_rows = pool_rows; a.pool_nrows = pool_nrows; a.HW = HW;
  a.style = ap.style; a.C = ap.C; a.R = ap.R; a.M = ap.M; a.A = ap.A; a.ca = ap.w[0];
  a.attributes = attributes; a.sq = sq; a.out_scale = out_scale; a.svec = svec; a.dyv = dyv; a.sig = sig;
  a.sig_stride = sig_stride;
  a.ymean = ymean;
  const int nchunk = a.nchunk;
  dim3 grid(nchunk, B);
  if (r_is_bf16)
    return launch_pdl(PDL_SIMT, bwd_reduce_ca_kernel<__nv_bfloat16>, grid, dim3(256), 0, s, g, reinterpret_cast<const __nv_bfloat16*>(r),
                      part, HW, C, nchunk, a, tickets) == cudaSuccess ? DFIR_OK : DFIR_ERR_CUDA;
  return launch_pdl(PDL_SIMT, bwd_reduce_ca_kernel<float>, grid, dim3(256), 0, s, g, reinterpret_cast<const float*>(r), part, HW, C,
                    nchunk, a, tickets) == cudaSuccess ? DFIR_OK : DFIR_ERR_CUDA;
}

int pa_backward_ctas(int B, int HW) {
  const long long nvec = static_cast<long long>(HW) * 8;
  return static_cast<int>(std::max<long long>(1, std::min<long long>((nvec + 2047) / 2048, (592 + B - 1) / B)));
}
size_t pa_backward_part_floats(int B, int HW) { return static_cast<size_t>(B) * pa_backward_ctas(B, HW) * kPaRec; }

int pa_backward(const float* g, const void* r, int r_is_bf16, const float* ymean, const AttnParams& ap,
                const float* attributes, const float* sq, const float* pa, float* du, float* part, int B, int HW,
                cudaStream_t s) {
  if (ap.C != 64 || ap.style == DFIR_STYLE_NONE || ap.A > 512 || ymean == nullptr || pa == nullptr) return DFIR_ERR_ARG;
  if (B <= 0 || HW <= 0) return DFIR_OK;
  dim3 grid(pa_backward_ctas(B, HW), B);
  if (r_is_bf16)
    pa_backward_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(g, reinterpret_cast<const __nv_bfloat16*>(r), ymean, ap, attributes,
                                                          sq, pa, du, part, HW);
  else
    pa_backward_kernel<float><<<grid, 256, 0, s>>>(g, reinterpret_cast<const float*>(r), ymean, ap, attributes, sq, pa, du,
                                                  part, HW);
  return ok_or_cuda3();
}

int pa_finish(const float* part, const float* sq, float out_scale, float* sig, int sig_stride, int dzq_off,
              float* const* gpa, int B, int HW, cudaStream_t s) {
  if (B <= 0 || HW <= 0) return DFIR_OK;
  pa_finish_kernel<<<(kPaRec + 127) / 128, 128, 0, s>>>(part, pa_backward_ctas(B, HW), B, sq, out_scale, sig, sig_stride,
                                                       dzq_off, gpa);
  return ok_or_cuda3();
}

int form_dr(const float* g, const float* svec, const float* dyv, void* dr, int dr_is_bf16, int B, int HW, int C,
            cudaStream_t s) {
  const long long nvec = static_cast<long long>(HW) * (C / 8);
  long long per_img = std::max<long long>(1, std::min<long long>((nvec + 2047) / 2048, (592 + B - 1) / B));
  dim3 grid(static_cast<unsigned>(per_img), B);
  if (dr_is_bf16)
    return launch_pdl(PDL_SIMT, form_dr_kernel<__nv_bfloat16>, grid, dim3(256), 0, s, g, svec, dyv, reinterpret_cast<__nv_bfloat16*>(dr),
                      HW, C) == cudaSuccess ? DFIR_OK : DFIR_ERR_CUDA;
  return launch_pdl(PDL_SIMT, form_dr_kernel<float>, grid, dim3(256), 0, s, g, svec, dyv, reinterpret_cast<float*>(dr), HW, C) ==
                 cudaSuccess ? DFIR_OK : DFIR_ERR_CUDA;
}

int add_f32(const float* a, const float* b, float* out, void* out_bf16, long long n, cudaStream_t s) {
  if (n % 4 != 0) return DFIR_ERR_ARG;
  if (n == 0) return DFIR_OK;
  return launch_pdl(PDL_SIMT, add_f32_kernel, dim3(grid_for(n / 4)), dim3(256), 0, s, reinterpret_cast<const float4*>(a),
                    reinterpret_cast<const float4*>(b), reinterpret_cast<float4*>(out), reinterpret_cast<uint2*>(out_bf16),
                    n / 4) == cudaSuccess ? DFIR_OK : DFIR_ERR_CUDA;
}

int adam_flat(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2, float eps, float wd,
              long long step, cudaStream_t s) {
  if (n % 4 != 0 || step < 1) return DFIR_ERR_ARG;
  if (n == 0) return DFIR_OK;
  const float bc1 = static_cast<float>(1.0 - pow(static_cast<double>(b1), static_cast<double>(step)));
  const float bc2s = static_cast<float>(sqrt(1.0 - pow(static_cast<double>(b2), static_cast<double>(step))));
  adam_flat_kernel<<<grid_for(n / 4), 256, 0, s>>>(reinterpret_cast<float4*>(p), reinterpret_cast<const float4*>(g),
                                                   reinterpret_cast<float4*>(m), reinterpret_cast<float4*>(v), n / 4, lr, b1,
                                                   b2, eps, wd, bc1, bc2s);
  return ok_or_cuda3();
}

int pixel_unshuffle_f32(const float* in, float* out, int B, int h, int w, int C, int r, cudaStream_t s) {
  const long long n = static_cast<long long>(B) * h * w * C * r * r;
  if (n == 0) return DFIR_OK;
  pixel_unshuffle_f32_kernel<<<grid_for(n), 256, 0, s>>>(in, out, B, h, w, C, r);
  return ok_or_cuda3();
}

int wgrad_f32_chunks(int B, int H, int Cin, int Cout) {
  const long long rows = static_cast<long long>(B) * H;
  const long long per = static_cast<long long>(9) * Cin * Cout * 4;
  long long S = std::min<long long>(rows, std::max<long long>(1, (48ll << 20) / per));
  const long long tiles = static_cast<long long>((Cin + 15) / 16) * ((Cout + 63) / 64);
  S = std::min<long long>(S, std::max<long long>(1, 592 / tiles));
  return static_cast<int>(std::max<long long>(1, S));
}

size_t wgrad_scratch_floats(int S, int Cin, int Cout) { return static_cast<size_t>(S) * (9ull * Cin * Cout + Cout); }

int wgrad_f32(const float* dY, const float* X, float* scratch, int B, int H, int W, int Cin, int Cout, cudaStream_t s,
              int* S_out) {
  const int S = wgrad_f32_chunks(B, H, Cin, Cout);
  *S_out = S;
  dim3 grid(S, (Cin + 15) / 16, (Cout + 63) / 64);
  float* dbpart = scratch + static_cast<size_t>(S) * 9 * Cin * Cout;
  wgrad_f32_kernel<<<grid, 256, 0, s>>>(dY, X, scratch, dbpart, B, H, W, Cin, Cout);
  return ok_or_cuda3();
}

int wgrad_reduce(const float* part, const float* dbpart, int S, int Cin, int n_rows, float* const* w_tbl, int w_idx,
                 float* w_direct, float* const* b_tbl, int b_idx, float* b_direct, int co_begin, int co_stride,
                 cudaStream_t s, int co_major) {
  const int total = 9 * Cin * n_rows + n_rows;
  return launch_pdl(PDL_WGRAD, wgrad_reduce_kernel, dim3((total + 63) / 64), dim3(256), 0, s, part, dbpart, S, Cin, n_rows, w_tbl,
                    w_idx, w_direct, b_tbl, b_idx, b_direct, co_begin, co_stride, co_major) == cudaSuccess ? DFIR_OK : DFIR_ERR_CUDA;
}

int wgrad_small_chunks(int B, int H) { return static_cast<int>(std::min<long long>(static_cast<long long>(B) * H, 296)); }
size_t wgrad_small_scratch_floats(int B, int H, int C) {
  return static_cast<size_t>(wgrad_small_chunks(B, H)) * (27ull * C + C + 4);
}

int wgrad_small(const float* I, const void* F, int f_is_bf16, float* scratch, int B, int H, int W, int C, int C3,
                int tail_mode, float* dw, float* db, cudaStream_t s) {
  if (C % 64 != 0 || C3 > 3 || C3 < 1) return DFIR_ERR_ARG;
  const int S = wgrad_small_chunks(B, H);
  float* kpart = scratch;
  float* fsum = kpart + static_cast<size_t>(S) * 27 * C;
  float* isum = fsum + static_cast<size_t>(S) * C;
  const size_t smem = (static_cast<size_t>(9) * (W + 2) + 4 * 28 * 64) * 4;
  if (smem > 200 * 1024) return DFIR_ERR_ARG;
  dim3 grid(S, C / 64);
  if (f_is_bf16) {
    auto k = wgrad_small_kernel<__nv_bfloat16>;
    if (smem > 48 * 1024 && cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess)
      return DFIR_ERR_CUDA;
    k<<<grid, 256, smem, s>>>(I, reinterpret_cast<const __nv_bfloat16*>(F), kpart, fsum, isum, B, H, W, C, C3);
  } else {
    auto k = wgrad_small_kernel<float>;
    if (smem > 48 * 1024 && cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess)
      return DFIR_ERR_CUDA;
    k<<<grid, 256, smem, s>>>(I, reinterpret_cast<const float*>(F), kpart, fsum, isum, B, H, W, C, C3);
  }
  if (cudaGetLastError() != cudaSuccess) return DFIR_ERR_CUDA;
  const int total = C3 * 9 * C;
  wgrad_small_reduce_kernel<<<(total + 255) / 256, 256, 0, s>>>(kpart, fsum, isum, S, C, C3, tail_mode, dw, db);
  return ok_or_cuda3();
}

int attn_param_grads(const float* sig, int sig_stride, const float* attributes, int A, const float* meta_w1,
                     const float* meta_b1, const float* meta_w2, const int* q_enabled, float* const* ca_g,
                     float* const* meta_g, int nblk, int B, int C, int R, int M, int Hid, int style, int meta_relu,
                     cudaStream_t s) {
  if (nblk <= 0) return DFIR_OK;
  AttnGradArgs a{};
  a.sig = sig; a.sig_stride = sig_stride; a.attributes = attributes; a.A = A; a.meta_w1 = meta_w1; a.meta_b1 = meta_b1;
  a.meta_w2 = meta_w2; a.q_enabled = q_enabled; a.ca_g = ca_g; a.meta_g = meta_g; a.B = B; a.C = C; a.R = R; a.M = M;
  a.Hid = Hid; a.style = style; a.meta_relu = meta_relu;
  const size_t smem = static_cast<size_t>(2) * B * std::max(1, Hid) * 4;
  if (smem > 160 * 1024) return DFIR_ERR_ARG;
  if (smem > 48 * 1024 &&
      cudaFuncSetAttribute(attn_param_grads_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess)
    return DFIR_ERR_CUDA;
  attn_param_grads_kernel<<<nblk, 256, smem, s>>>(a);
  return ok_or_cuda3();
}

}  // namespace dfir
