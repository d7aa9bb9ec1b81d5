// CUDA-core kernels of the Deep-FIR hot path: weight packing, head conv, the fp32 parity-mode conv, the
// attention-vector kernels and the fused channel-scale + residual streamer.  These are the HBM / latency
// bound pieces: they are written for coalesced 16-byte accesses, not tensor cores.
#include "kernels.h"
#include "attn.cuh"

#include <algorithm>
#include "launch.cuh"
#include "ptx.cuh"

namespace dfir {

namespace {

// ------------------------------------------------------------------------------------------------
// weight packing
// ------------------------------------------------------------------------------------------------
__global__ void pack_conv_bf16_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int cout, int cin,
                                      int nt_rows, int co_begin, int co_stride) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int total = 9 * nt_rows * 64;
  if (idx >= total) return;
  const int ci = idx % 64;
  const int n = (idx / 64) % nt_rows;
  const int t = idx / (64 * nt_rows);
  const int co = co_begin + n * co_stride;
  float v = 0.f;
  if (co < cout && ci < cin) v = w[(static_cast<size_t>(co) * cin + ci) * 9 + t];
  // K-major SWIZZLE_128B: row n = 128 B, 16-byte chunk index XORed with (n & 7)
  const int chunk = (ci >> 3) ^ (n & 7);
  out[static_cast<size_t>(t) * nt_rows * 64 + n * 64 + chunk * 8 + (ci & 7)] = __float2bfloat16_rn(v);
}

__global__ void pack_conv_f32_kernel(const float* __restrict__ w, float* __restrict__ out, int cout, int cin) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int total = 9 * cin * cout;
  if (idx >= total) return;
  const int co = idx % cout;
  const int ci = (idx / cout) % cin;
  const int t = idx / (cout * cin);
  out[idx] = w[(static_cast<size_t>(co) * cin + ci) * 9 + t];
}

// ------------------------------------------------------------------------------------------------
// head conv: NCHW fp32 image (Cin small) -> NHWC features.  8 threads per pixel, 8 output channels each.
// ------------------------------------------------------------------------------------------------
// One item = 128 consecutive pixels of one image row x all Cout channels.  256 threads = 8 warps; warp = one group of 8
// output channels (wider nets loop over further groups), lane = 4 consecutive pixels: 32 accumulators per thread, so one
// pair of broadcast weight loads (8 floats) and two patch loads (8 floats) feed 32 / 96 FMAs — the first versions issued
// three shared-memory loads per 8 FMAs and ran at 12-16 % of HBM bandwidth.  The 3 x 130 x Cin input patch is staged in
// shared memory (coalesced row reads of the NCHW planes), the 9 * Cin * Cout weights once per CTA; items are
// grid-strided with 32-bit index arithmetic.
constexpr int kHeadPatchStride = 136;  // floats per patch row: 130 used, 16-byte aligned rows
__global__ void __launch_bounds__(256)
head_conv_kernel(const float* __restrict__ x, const float* __restrict__ wp, const float* __restrict__ bias,
                 float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_bf16, int B, int Cin, int H, int W, int Cout,
                 __nv_bfloat16* __restrict__ out_lo, int lo8) {
  extern __shared__ __align__(16) float ws[];  // [9*Cin][Cout] weights, then [Cin][3][kHeadPatchStride] input patch
  const int nw = 9 * Cin * Cout;
  float* patch = ws + ((nw + 3) & ~3);
  for (int i = threadIdx.x; i < nw; i += blockDim.x) ws[i] = wp[i];
  const int oct_per_pix = Cout / 8;
  const int xchunks = (W + 127) / 128;
  const int nitems = B * H * xchunks;
  const int lane = threadIdx.x & 31, og = threadIdx.x >> 5;
  for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
    const int xc = item % xchunks, row = item / xchunks;
    const int y = row % H, b = row / H;
    const int x0 = xc * 128;
    __syncthreads();  // previous item's patch reads done (also orders the weight staging before the first use)
    for (int i = threadIdx.x; i < Cin * 3 * kHeadPatchStride; i += blockDim.x) {
      const int dx = i % kHeadPatchStride, r = i / kHeadPatchStride;
      const int dy = r % 3, ci = r / 3;
      const int yy = y + dy - 1, xx = x0 + dx - 1;
      patch[i] = (dx < 130 && yy >= 0 && yy < H && xx >= 0 && xx < W)
                     ? x[((static_cast<size_t>(b) * Cin + ci) * H + yy) * W + xx] : 0.f;
    }
    __syncthreads();
    const int xw = x0 + 4 * lane;  // first of this thread's four pixels
    if (xw >= W) continue;
    for (int oct = og; oct < oct_per_pix; oct += 8) {
      const int oc = oct * 8;
      float acc[4][8];
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[p][i] = bias != nullptr ? bias[oc + i] : 0.f;
      for (int ci = 0; ci < Cin; ++ci)
#pragma unroll
        for (int dy = 0; dy < 3; ++dy) {
          const float4* pr = reinterpret_cast<const float4*>(patch + (ci * 3 + dy) * kHeadPatchStride + 4 * lane);
          const float4 p0 = pr[0], p1 = pr[1];
          const float v[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) {
            const float4* wr = reinterpret_cast<const float4*>(ws + ((dy * 3 + dx) * Cin + ci) * Cout + oc);
            const float4 w0 = wr[0], w1 = wr[1];
            const float w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
              for (int i = 0; i < 8; ++i) acc[p][i] = fmaf(v[p + dx], w[i], acc[p][i]);
          }
        }
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        if (xw + p >= W) break;
        const size_t o = ((static_cast<size_t>(b) * H + y) * W + xw + p) * Cout + oc;
        if (out_f32 != nullptr) {
          *reinterpret_cast<float4*>(out_f32 + o) = make_float4(acc[p][0], acc[p][1], acc[p][2], acc[p][3]);
          *reinterpret_cast<float4*>(out_f32 + o + 4) = make_float4(acc[p][4], acc[p][5], acc[p][6], acc[p][7]);
        }
        if (out_bf16 != nullptr) {
          __align__(16) __nv_bfloat162 pk[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) pk[i] = __floats2bfloat162_rn(acc[p][2 * i], acc[p][2 * i + 1]);
          *reinterpret_cast<uint4*>(out_bf16 + o) = *reinterpret_cast<uint4*>(pk);
          if (out_lo != nullptr && lo8) {
            // 8-bit lo plane (see EPI_SCALE_SKIP_HL8 in conv_tc.cu): the stream value is the 24-bit float X nearest to the
            // result, bits(X) = (hi << 16) + (q << 8); hi = nearest bf16 with ties away from zero REPLACES the RN-even hi
            __align__(16) unsigned short hb[8];
            unsigned char qb[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const uint32_t t = __float_as_uint(acc[p][i]) + 0x80u;
              const uint32_t bh = (t + 0x8000u) & 0xffff0000u;
              hb[i] = static_cast<unsigned short>(bh >> 16);
              qb[i] = static_cast<unsigned char>(((t - bh) >> 8) & 0xffu);
            }
            *reinterpret_cast<uint4*>(out_bf16 + o) = *reinterpret_cast<uint4*>(hb);
            // the lo plane keeps a pixel's 64 bytes in accumulator-fragment order: byte cq * 16 + 2 n + e = channel 8 n + 2 cq + e
            // (this thread: n = oc / 8, channels oc + 2 cq + e)
            unsigned char* lo_px = reinterpret_cast<unsigned char*>(out_lo) + (o - oc) + 2 * (oc >> 3);
#pragma unroll
            for (int cqi = 0; cqi < 4; ++cqi)
              *reinterpret_cast<unsigned short*>(lo_px + 16 * cqi) =
                  static_cast<unsigned short>(qb[2 * cqi] | (static_cast<unsigned short>(qb[2 * cqi + 1]) << 8));
          } else if (out_lo != nullptr) {  // hi + lo residual stream: lo = bf16(value - hi)
            __align__(16) __nv_bfloat162 lo[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 h = __bfloat1622float2(pk[i]);
              lo[i] = __floats2bfloat162_rn(acc[p][2 * i] - h.x, acc[p][2 * i + 1] - h.y);
            }
            *reinterpret_cast<uint4*>(out_lo + o) = *reinterpret_cast<uint4*>(lo);
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// fp32 parity-mode conv (CUDA cores): NHWC fp32, each thread = 4 consecutive pixels of a row x 8 output
// channels.  Per (dy, 4 cin): 6 float4 input loads + 24 float4 weight loads feed 384 FMAs.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
conv3x3_f32_kernel(const float* __restrict__ in, const float* __restrict__ wp, const float* __restrict__ bias,
                   const float* __restrict__ skip, float* __restrict__ out, int B, int H, int W, int Cin, int Cout,
                   int relu, int ps_r, int out_nchw, const float* __restrict__ mask) {
  const int n_oct = (Cout + 7) / 8;
  const int quads_per_row = (W + 3) / 4;
  const long long nquad = static_cast<long long>(B) * H * quads_per_row;
  const long long gid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long quad = gid / n_oct;
  const int oc = static_cast<int>(gid % n_oct) * 8;
  if (quad >= nquad) return;
  const int x0 = static_cast<int>(quad % quads_per_row) * 4;
  const int y = static_cast<int>((quad / quads_per_row) % H);
  const int b = static_cast<int>(quad / (static_cast<long long>(quads_per_row) * H));
  const bool full_oct = (oc + 8 <= Cout) && (Cout % 4 == 0);

  float acc[4][8];
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[p][i] = (bias != nullptr && oc + i < Cout) ? bias[oc + i] : 0.f;

  for (int dy = 0; dy < 3; ++dy) {
    const int yy = y + dy - 1;
    if (yy < 0 || yy >= H) continue;
    const float* inrow = in + (static_cast<size_t>(b) * H + yy) * W * Cin;
    for (int ci = 0; ci < Cin; ci += 4) {
      float4 v[6];
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        const int xx = x0 + j - 1;
        v[j] = (xx >= 0 && xx < W) ? *reinterpret_cast<const float4*>(inrow + static_cast<size_t>(xx) * Cin + ci)
                                   : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          const float* wr = wp + (static_cast<size_t>(dy * 3 + dx) * Cin + ci + cc) * Cout + oc;
          float w8[8];
          if (full_oct) {
            const float4 w0 = *reinterpret_cast<const float4*>(wr);
            const float4 w1 = *reinterpret_cast<const float4*>(wr + 4);
            w8[0] = w0.x; w8[1] = w0.y; w8[2] = w0.z; w8[3] = w0.w;
            w8[4] = w1.x; w8[5] = w1.y; w8[6] = w1.z; w8[7] = w1.w;
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) w8[i] = (oc + i < Cout) ? wr[i] : 0.f;
          }
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            const float4 t = v[p + dx];
            const float a = cc == 0 ? t.x : (cc == 1 ? t.y : (cc == 2 ? t.z : t.w));
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[p][i] = fmaf(a, w8[i], acc[p][i]);
          }
        }
      }
    }
  }

#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const int x = x0 + p;
    if (x >= W) continue;
    const size_t pix = (static_cast<size_t>(b) * H + y) * W + x;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = oc + i;
      if (k >= Cout) continue;
      float val = acc[p][i];
      if (skip != nullptr) val += skip[pix * Cout + k];
      if (relu) val = fmaxf(val, 0.f);
      if (mask != nullptr && !(mask[pix * Cout + k] > 0.f)) val = 0.f;  // ReLU backward against the saved activation
      size_t o;
      if (out_nchw) {
        o = ((static_cast<size_t>(b) * Cout + k) * H + y) * W + x;
      } else if (ps_r > 1) {
        const int rr = ps_r * ps_r;
        const int c = k / rr, ij = k % rr, ii = ij / ps_r, jj = ij % ps_r;
        const int Co = Cout / rr;
        o = ((static_cast<size_t>(b) * H * ps_r + (y * ps_r + ii)) * (static_cast<size_t>(W) * ps_r) + (x * ps_r + jj)) * Co + c;
      } else {
        o = pix * Cout + k;
      }
      out[o] = val;
    }
  }
}

// per-row channel sums, fp32 NHWC: one block per (b, y), thread c sums over x
__global__ void pool_rows_f32_kernel(const float* __restrict__ in, float* __restrict__ pool_rows, int W, int C) {
  const size_t row = blockIdx.x;
  const float* p = in + row * W * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s = 0.f;
    for (int x = 0; x < W; ++x) s += p[static_cast<size_t>(x) * C + c];
    pool_rows[row * C + c] = s;
  }
}

// ------------------------------------------------------------------------------------------------
// meta-attention vectors for all blocks in one launch: grid (nblk, B), block = C threads
// ------------------------------------------------------------------------------------------------
__global__ void meta_attention_kernel(const float* __restrict__ meta, const float* __restrict__ w1,
                                      const float* __restrict__ b1, const float* __restrict__ w2,
                                      const float* __restrict__ b2, float* __restrict__ out, int B, int M, int Hid,
                                      int C, int relu, const int* __restrict__ blk_enabled, float out_scale) {
  extern __shared__ float sh[];  // [M] meta, [Hid] hidden
  float* m_s = sh;
  float* h_s = sh + M;
  const int blk = blockIdx.x;
  const int b = blockIdx.y;
  float* o = out + (static_cast<size_t>(blk) * B + b) * C;
  if (blk_enabled != nullptr && blk_enabled[blk] == 0) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) o[c] = out_scale;
    return;
  }
  for (int i = threadIdx.x; i < M; i += blockDim.x) m_s[i] = meta[static_cast<size_t>(b) * M + i];
  __syncthreads();
  for (int h = threadIdx.x; h < Hid; h += blockDim.x) {
    const float* wr = w1 + (static_cast<size_t>(blk) * Hid + h) * M;
    float s = b1[static_cast<size_t>(blk) * Hid + h];
    for (int i = 0; i < M; ++i) s = fmaf(wr[i], m_s[i], s);
    h_s[h] = relu ? fmaxf(s, 0.f) : s;
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float* wr = w2 + (static_cast<size_t>(blk) * C + c) * Hid;
    float s = b2[static_cast<size_t>(blk) * C + c];
    for (int h = 0; h < Hid; ++h) s = fmaf(wr[h], h_s[h], s);
    o[c] = out_scale / (1.f + expf(-s));
  }
}

// ------------------------------------------------------------------------------------------------
// fused: pool finalise -> attention vector -> x_out = r * s + x_in (fp32) (+ bf16 copy)
// grid (ctas_per_image, B), 256 threads; each thread streams 8 channels of a pixel per iteration.
// ------------------------------------------------------------------------------------------------
template <bool R_BF16>
__global__ void __launch_bounds__(256)
scale_residual_kernel(const void* __restrict__ r_, const float* __restrict__ x_in,
                      const float* __restrict__ pool_rows, int pool_nrows, AttnParams ap,
                      const float* __restrict__ attributes, const float* __restrict__ sq, float res_scale,
                      float* __restrict__ x_out, __nv_bfloat16* __restrict__ x_out_bf16, int HW, int C,
                      float* __restrict__ y_out, const float* __restrict__ pa) {
  __shared__ float y_s[256];
  __shared__ float s_s[256];
  __shared__ float attr_s[512];
  __shared__ float tmp[4 * 256];
  __shared__ float pa_s[8 * 64 + 8 + 8 + 4];  // PALayer: W1[8][64] b1[8] W2[8] b2 (C == 64 only)
  __shared__ float caw_s[1024];               // this block's attention-MLP parameters (when they fit)
  const int b = blockIdx.y;
  const int tid = threadIdx.x;
  // constants first: under programmatic dependent launch this part overlaps the tail of the previous kernel
  if (pa != nullptr)
    for (int i = tid; i < 8 * 64 + 17; i += 256) pa_s[i] = pa[i];
  const int n_caw = attn_param_count(ap.style, C, ap.R, ap.M);
  const float* ca_w = ap.w[0];
  if (n_caw > 0 && n_caw <= 1024) {
    for (int i = tid; i < n_caw; i += 256) caw_s[i] = ap.w[0][i];
    ca_w = caw_s;
  }
  ptx::grid_dep_wait();  // (no early launch_dependents: a resident, waiting successor would starve our later waves)

  if (ap.style != DFIR_STYLE_NONE) {
    // deterministic pooled mean: 256/C row groups, fixed-order combine
    const int ngrp = 256 / C;
    const int c = tid % C, grp = tid / C;
    float s = 0.f;
    if (grp < ngrp) {
      const float* pr = pool_rows + static_cast<size_t>(b) * pool_nrows * C + c;
      for (int row = grp; row < pool_nrows; row += ngrp) s += pr[static_cast<size_t>(row) * C];
      tmp[grp * C + c] = s;
    }
    for (int i = tid; i < ap.A; i += 256) attr_s[i] = attributes[static_cast<size_t>(b) * ap.A + i];
    __syncthreads();
    if (tid < C) {
      float t = 0.f;
      for (int gI = 0; gI < ngrp; ++gI) t += tmp[gI * C + tid];
      y_s[tid] = t / static_cast<float>(HW);
      if (y_out != nullptr && blockIdx.x == 0) y_out[static_cast<size_t>(b) * C + tid] = y_s[tid];  // saved for backward
    }
    __syncthreads();
    attn_vector(BlockGroup{}, ap.style, ca_w, C, ap.R, ap.M, attr_s, y_s, s_s, tmp);
    // with pixel attention the meta scale is applied after it (QRCAB.forward, architectures.py:174-178)
    if (tid < C && pa == nullptr) s_s[tid] *= (sq != nullptr ? sq[static_cast<size_t>(b) * C + tid] : 1.f);
  } else {
    if (tid < C) s_s[tid] = res_scale * (sq != nullptr ? sq[static_cast<size_t>(b) * C + tid] : 1.f);
  }
  __syncthreads();

  const int lanes_per_pix = C / 8;
  const int c0 = (tid % lanes_per_pix) * 8;
  float sc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) sc[i] = s_s[c0 + i];

  const size_t img_off = static_cast<size_t>(b) * HW * C;
  const long long nvec = static_cast<long long>(HW) * lanes_per_pix;  // 8-channel vectors in this image
  if (pa != nullptr) {
    // PALayer (architectures.py:13-26): r <- r * sigmoid(W2 relu(W1 r + b1) + b2), one scalar per pixel, between the
    // channel attention and the meta attention.  The 8 threads of a pixel each hold 8 channels: partial 64 -> 8
    // products, butterfly over the 8 lanes, then every lane evaluates the 8 -> 1 layer.
    float sqv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) sqv[i] = (ap.style != DFIR_STYLE_NONE && sq != nullptr) ? sq[static_cast<size_t>(b) * C + c0 + i] : 1.f;
    for (long long base = static_cast<long long>(blockIdx.x) * 256; base < nvec;
         base += static_cast<long long>(gridDim.x) * 256) {
      const long long vi = base + tid;
      const bool active = vi < nvec;
      const size_t e = img_off + static_cast<size_t>(active ? vi : 0) * 8;
      float u[8];
      if (R_BF16) {
        const uint4 raw = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(r_) + e);
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 f = __bfloat1622float2(h[i]);
          u[2 * i] = f.x;
          u[2 * i + 1] = f.y;
        }
      } else {
        const float4 a0 = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(r_) + e);
        const float4 a1 = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(r_) + e + 4);
        u[0] = a0.x; u[1] = a0.y; u[2] = a0.z; u[3] = a0.w; u[4] = a1.x; u[5] = a1.y; u[6] = a1.z; u[7] = a1.w;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) u[i] *= sc[i];
      float hsum[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t = fmaf(pa_s[j * 64 + c0 + i], u[i], t);
        hsum[j] = t;
      }
#pragma unroll
      for (int mask = 1; mask < 8; mask <<= 1)
#pragma unroll
        for (int j = 0; j < 8; ++j) hsum[j] += __shfl_xor_sync(0xffffffffu, hsum[j], mask);
      float z = pa_s[8 * 64 + 16];
#pragma unroll
      for (int j = 0; j < 8; ++j) z = fmaf(pa_s[8 * 64 + 8 + j], fmaxf(hsum[j] + pa_s[8 * 64 + j], 0.f), z);
      const float pmap = 1.f / (1.f + expf(-z));
      if (active) {
        const float4 x0 = *reinterpret_cast<const float4*>(x_in + e);
        const float4 x1 = *reinterpret_cast<const float4*>(x_in + e + 4);
        const float xi[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = fmaf(u[i] * pmap, sqv[i], xi[i]);
        *reinterpret_cast<float4*>(x_out + e) = make_float4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<float4*>(x_out + e + 4) = make_float4(o[4], o[5], o[6], o[7]);
        if (x_out_bf16 != nullptr) {
          __align__(16) __nv_bfloat162 pk[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) pk[i] = __floats2bfloat162_rn(o[2 * i], o[2 * i + 1]);
          *reinterpret_cast<uint4*>(x_out_bf16 + e) = *reinterpret_cast<uint4*>(pk);
        }
      }
    }
    return;
  }
  for (long long vi = static_cast<long long>(blockIdx.x) * 256 + tid; vi < nvec;
       vi += static_cast<long long>(gridDim.x) * 256) {
    const size_t e = img_off + static_cast<size_t>(vi) * 8;
    float rv[8];
    if (R_BF16) {
      const uint4 raw = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(r_) + e);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 f = __bfloat1622float2(h[i]);
        rv[2 * i] = f.x;
        rv[2 * i + 1] = f.y;
      }
    } else {
      const float4 a0 = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(r_) + e);
      const float4 a1 = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(r_) + e + 4);
      rv[0] = a0.x; rv[1] = a0.y; rv[2] = a0.z; rv[3] = a0.w;
      rv[4] = a1.x; rv[5] = a1.y; rv[6] = a1.z; rv[7] = a1.w;
    }
    const float4 x0 = *reinterpret_cast<const float4*>(x_in + e);
    const float4 x1 = *reinterpret_cast<const float4*>(x_in + e + 4);
    float o[8];
    o[0] = fmaf(rv[0], sc[0], x0.x); o[1] = fmaf(rv[1], sc[1], x0.y);
    o[2] = fmaf(rv[2], sc[2], x0.z); o[3] = fmaf(rv[3], sc[3], x0.w);
    o[4] = fmaf(rv[4], sc[4], x1.x); o[5] = fmaf(rv[5], sc[5], x1.y);
    o[6] = fmaf(rv[6], sc[6], x1.z); o[7] = fmaf(rv[7], sc[7], x1.w);
    *reinterpret_cast<float4*>(x_out + e) = make_float4(o[0], o[1], o[2], o[3]);
    *reinterpret_cast<float4*>(x_out + e + 4) = make_float4(o[4], o[5], o[6], o[7]);
    if (x_out_bf16 != nullptr) {
      __align__(16) __nv_bfloat162 pk[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) pk[i] = __floats2bfloat162_rn(o[2 * i], o[2 * i + 1]);
      *reinterpret_cast<uint4*>(x_out_bf16 + e) = *reinterpret_cast<uint4*>(pk);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Channel attention of an RCAB from statistics of t = relu(conv1(x)), i.e. BEFORE conv2 runs.
//   mean over pixels of r = conv2(t) is linear in t:
//     mean(r)[co] = b2[co] + (1/HW) * sum_{tap,ci} W2[co][ci][tap] * S[tap][ci],
//     S[tap][ci]  = sum over the pixels of t[ci] that the tap (dy,dx) reads inside the image
//                 = T - (row excluded by dy) - (column excluded by dx) + (their corner)
//   with T = total sum, first/last row sums and first/last column sums of t.  conv2's epilogue can then apply
//   the attention scale and the residual add itself, and r never goes to memory.
// grid = B images, 256 threads.  All reductions run in a fixed order (bit-reproducible).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
ca_from_stats_kernel(const float* __restrict__ pool_rows, const float* __restrict__ col_first,
                     const float* __restrict__ col_last, const __nv_bfloat16* __restrict__ w2,
                     const float* __restrict__ bias2, AttnParams ap, const float* __restrict__ attributes,
                     const float* __restrict__ sq, float* __restrict__ svec, int H, int W, int nseg) {
  __shared__ float part[4][5][64];  // partial sums: T, R0, RL, C0, CL by 4 row groups
  __shared__ float S[9][64];
  __shared__ float y_s[64];
  __shared__ float s_s[64];
  __shared__ float attr_s[512];
  __shared__ float tmp[1024];
  const int b = blockIdx.x;
  const int tid = threadIdx.x;
  const int c = tid & 63, grp = tid >> 6;
  const float* pr = pool_rows + static_cast<size_t>(b) * nseg * H * 64;
  const float* cf = col_first + static_cast<size_t>(b) * H * 64;
  const float* cl = col_last + static_cast<size_t>(b) * H * 64;
  {
    float t = 0.f, c0 = 0.f, c1 = 0.f;
    for (int seg = 0; seg < nseg; ++seg)
      for (int y = grp; y < H; y += 4) t += pr[(static_cast<size_t>(seg) * H + y) * 64 + c];
    for (int y = grp; y < H; y += 4) {
      c0 += cf[static_cast<size_t>(y) * 64 + c];
      c1 += cl[static_cast<size_t>(y) * 64 + c];
    }
    float r0 = 0.f, rl = 0.f;
    if (grp == 0)
      for (int seg = 0; seg < nseg; ++seg) {
        r0 += pr[(static_cast<size_t>(seg) * H + 0) * 64 + c];
        rl += pr[(static_cast<size_t>(seg) * H + (H - 1)) * 64 + c];
      }
    part[grp][0][c] = t; part[grp][1][c] = r0; part[grp][2][c] = rl; part[grp][3][c] = c0; part[grp][4][c] = c1;
  }
  for (int i = tid; i < ap.A; i += 256) attr_s[i] = attributes[static_cast<size_t>(b) * ap.A + i];
  __syncthreads();
  if (tid < 64) {
    const float T = ((part[0][0][c] + part[1][0][c]) + part[2][0][c]) + part[3][0][c];
    const float R0 = part[0][1][c], RL = part[0][2][c];
    const float C0 = ((part[0][3][c] + part[1][3][c]) + part[2][3][c]) + part[3][3][c];
    const float CL = ((part[0][4][c] + part[1][4][c]) + part[2][4][c]) + part[3][4][c];
    const float k00 = cf[c], k0w = cl[c];                                  // t[0][0], t[0][W-1]
    const float kh0 = cf[static_cast<size_t>(H - 1) * 64 + c], khw = cl[static_cast<size_t>(H - 1) * 64 + c];
#pragma unroll
    for (int dy = 0; dy < 3; ++dy)
#pragma unroll
      for (int dx = 0; dx < 3; ++dx) {
        // tap offset (oy,ox) = (dy-1,dx-1): oy=-1 never reads the last row, oy=+1 never the first row (same for x)
        const float rowx = dy == 0 ? RL : (dy == 2 ? R0 : 0.f);
        const float colx = dx == 0 ? CL : (dx == 2 ? C0 : 0.f);
        float corner = 0.f;
        if (dy == 0 && dx == 0) corner = khw;
        if (dy == 0 && dx == 2) corner = kh0;
        if (dy == 2 && dx == 0) corner = k0w;
        if (dy == 2 && dx == 2) corner = k00;
        S[dy * 3 + dx][c] = T - rowx - colx + corner;
      }
  }
  __syncthreads();
  {
    // y[co] = b2[co] + (1/HW) * sum W2[co][:][tap] . S[tap][:]; 4 threads per output channel split the taps.
    const int co = tid >> 2, part4 = tid & 3;
    float acc = 0.f;
    for (int tap = part4; tap < 9; tap += 4) {
      const __nv_bfloat16* wr = w2 + (static_cast<size_t>(tap) * 64 + co) * 64;  // one swizzled 128-byte row
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        const uint4 raw = *reinterpret_cast<const uint4*>(wr + ((ch ^ (co & 7)) << 3));
        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __bfloat1622float2(h2[j]);
          acc = fmaf(f.x, S[tap][ch * 8 + 2 * j], acc);
          acc = fmaf(f.y, S[tap][ch * 8 + 2 * j + 1], acc);
        }
      }
    }
    tmp[tid] = acc;
  }
  __syncthreads();
  if (tid < 64) {
    const float tot = ((tmp[4 * tid] + tmp[4 * tid + 1]) + tmp[4 * tid + 2]) + tmp[4 * tid + 3];
    y_s[tid] = bias2[tid] + tot / (static_cast<float>(H) * static_cast<float>(W));
  }
  __syncthreads();
  attn_vector(BlockGroup{}, ap.style, ap.w[0], 64, ap.R, ap.M, attr_s, y_s, s_s, tmp);
  if (tid < 64) svec[static_cast<size_t>(b) * 64 + tid] = s_s[tid] * (sq != nullptr ? sq[static_cast<size_t>(b) * 64 + tid] : 1.f);
}

// ------------------------------------------------------------------------------------------------
// Post-processing of the SR batch on the device (ModelInterface.net_run_and_process,
// /root/reference/Code/SISR/models/__init__.py:138-169): rgb = clip(x, 0, 1); ycbcr = ycbcr_convert(rgb, im_type='jpg')
// (sr_tools/image_manipulation.py:56-157).  Every operation is an individually rounded fp32 multiply / add in the
// order numpy evaluates the reference expression, so the result is bit-identical to the CPU path.  HBM-bound:
// 12 B read + 24 B written per pixel.
// ------------------------------------------------------------------------------------------------
__global__ void postprocess_rgb_kernel(const float* __restrict__ x, float* __restrict__ rgb, float* __restrict__ ycc,
                                       long long HW, float lo, float hi, float bias_c) {
  const long long b = blockIdx.y;
  const float* xr = x + b * 3 * HW;
  float* orgb = rgb + b * 3 * HW;
  float* oy = ycc + b * 3 * HW;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < HW;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float r = fminf(fmaxf(xr[i], lo), hi), g = fminf(fmaxf(xr[HW + i], lo), hi), bl = fminf(fmaxf(xr[2 * HW + i], lo), hi);
    orgb[i] = r; orgb[HW + i] = g; orgb[2 * HW + i] = bl;
    const float y = __fadd_rn(__fadd_rn(__fmul_rn(0.299f, r), __fmul_rn(0.587f, g)), __fmul_rn(0.114f, bl));
    const float cb = __fadd_rn(bias_c, __fadd_rn(__fsub_rn(__fmul_rn(-0.168736f, r), __fmul_rn(0.331264f, g)), __fmul_rn(0.5f, bl)));
    const float cr = __fadd_rn(bias_c, __fsub_rn(__fsub_rn(__fmul_rn(0.5f, r), __fmul_rn(0.418688f, g)), __fmul_rn(0.081312f, bl)));
    oy[i] = y; oy[HW + i] = cb; oy[2 * HW + i] = cr;
  }
}

// The same post-processing with 8-bit results (what the evaluation loop writes to disk) and, when a ground-truth batch is
// given, the squared Y-channel error of every image (Metrics.run_image_metric('PSNR') compares channel 0 of the YCbCr
// arrays, sr_tools/metrics.py:6-17): 12 (+12) B read and 6 B written per pixel instead of 24 B written and a numpy pass.
// Partial sums are per block and fp64; the finishing kernel adds them in block order (deterministic).
__device__ __forceinline__ float ycc_y(float r, float g, float bl) {
  return __fadd_rn(__fadd_rn(__fmul_rn(0.299f, r), __fmul_rn(0.587f, g)), __fmul_rn(0.114f, bl));
}
__global__ void __launch_bounds__(256)
postprocess_u8_kernel(const float* __restrict__ x, const float* __restrict__ hr, unsigned char* __restrict__ rgb8,
                      unsigned char* __restrict__ ycc8, double* __restrict__ part, long long HW, float bias_c) {
  const long long b = blockIdx.y;
  const float* xr = x + b * 3 * HW;
  const float* hp = hr != nullptr ? hr + b * 3 * HW : nullptr;
  double acc = 0.0;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < HW;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float r = fminf(fmaxf(xr[i], 0.f), 1.f), g = fminf(fmaxf(xr[HW + i], 0.f), 1.f), bl = fminf(fmaxf(xr[2 * HW + i], 0.f), 1.f);
    const float y = ycc_y(r, g, bl);
    if (rgb8 != nullptr) {
      unsigned char* o = rgb8 + b * 3 * HW;
      o[i] = static_cast<unsigned char>(__float2int_rn(r * 255.f));
      o[HW + i] = static_cast<unsigned char>(__float2int_rn(g * 255.f));
      o[2 * HW + i] = static_cast<unsigned char>(__float2int_rn(bl * 255.f));
    }
    if (ycc8 != nullptr) {
      const float cb = __fadd_rn(bias_c, __fadd_rn(__fsub_rn(__fmul_rn(-0.168736f, r), __fmul_rn(0.331264f, g)), __fmul_rn(0.5f, bl)));
      const float cr = __fadd_rn(bias_c, __fsub_rn(__fsub_rn(__fmul_rn(0.5f, r), __fmul_rn(0.418688f, g)), __fmul_rn(0.081312f, bl)));
      unsigned char* o = ycc8 + b * 3 * HW;
      o[i] = static_cast<unsigned char>(__float2int_rn(fminf(fmaxf(y, 0.f), 1.f) * 255.f));
      o[HW + i] = static_cast<unsigned char>(__float2int_rn(fminf(fmaxf(cb, 0.f), 1.f) * 255.f));
      o[2 * HW + i] = static_cast<unsigned char>(__float2int_rn(fminf(fmaxf(cr, 0.f), 1.f) * 255.f));
    }
    if (hp != nullptr) {
      const float hr_ = fminf(fmaxf(hp[i], 0.f), 1.f), hg = fminf(fmaxf(hp[HW + i], 0.f), 1.f), hb = fminf(fmaxf(hp[2 * HW + i], 0.f), 1.f);
      const float d = __fsub_rn(y, ycc_y(hr_, hg, hb));
      acc += static_cast<double>(__fmul_rn(d, d));
    }
  }
  if (part != nullptr) {
    __shared__ double red[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int w = 0; w < 8; ++w) t += red[w];
      part[b * gridDim.x + blockIdx.x] = t;
    }
  }
}
__global__ void y_psnr_finish_kernel(const double* __restrict__ part, int nparts, long long HW, float* __restrict__ psnr) {
  const int b = blockIdx.x;
  if (threadIdx.x != 0) return;
  double t = 0.0;
  for (int i = 0; i < nparts; ++i) t += part[static_cast<long long>(b) * nparts + i];
  const double mse = t / static_cast<double>(HW);
  psnr[b] = mse == 0.0 ? 100.f : static_cast<float>(20.0 * log10(1.0 / sqrt(mse)));  // metrics.py:14-17, max_value = 1
}

inline int ok_or_cuda() { return cudaGetLastError() == cudaSuccess ? DFIR_OK : DFIR_ERR_CUDA; }

}  // namespace

int postprocess_u8(const float* x, const float* hr, unsigned char* rgb8, unsigned char* ycc8, float* y_psnr, double* scratch,
                   int B, long long HW, cudaStream_t s) {
  if (B <= 0 || HW <= 0) return DFIR_OK;
  if ((hr != nullptr) != (y_psnr != nullptr) || (hr != nullptr && scratch == nullptr)) return DFIR_ERR_ARG;
  const int nb = postprocess_u8_blocks(B, HW);
  dim3 grid(nb, B);
  postprocess_u8_kernel<<<grid, 256, 0, s>>>(x, hr, rgb8, ycc8, hr != nullptr ? scratch : nullptr, HW,
                                             static_cast<float>(128. * (1. / 255)));
  if (hr != nullptr) y_psnr_finish_kernel<<<B, 32, 0, s>>>(scratch, nb, HW, y_psnr);
  return ok_or_cuda();
}
int postprocess_u8_blocks(int B, long long HW) {
  return static_cast<int>(std::min<long long>((HW + 255) / 256, 1184 / std::max(1, std::min(B, 8)) + 1));
}

int pack_conv_weights_bf16(const float* w, void* out, int cout, int cin, int nt_rows, int co_begin, int co_stride,
                           cudaStream_t s) {
  if (cin > 64 || (nt_rows != 64 && nt_rows != 16)) return DFIR_ERR_ARG;
  const int total = 9 * nt_rows * 64;
  pack_conv_bf16_kernel<<<(total + 255) / 256, 256, 0, s>>>(w, reinterpret_cast<__nv_bfloat16*>(out), cout, cin,
                                                            nt_rows, co_begin, co_stride);
  return ok_or_cuda();
}

int pack_conv_weights_f32(const float* w, float* out, int cout, int cin, cudaStream_t s) {
  const int total = 9 * cin * cout;
  pack_conv_f32_kernel<<<(total + 255) / 256, 256, 0, s>>>(w, out, cout, cin);
  return ok_or_cuda();
}

int head_conv(const float* x, const float* wp, const float* bias, float* out_f32, __nv_bfloat16* out_bf16, int B,
              int Cin, int H, int W, int Cout, cudaStream_t s, __nv_bfloat16* out_lo, int lo8) {
  const int smem = (((9 * Cin * Cout + 3) & ~3) + Cin * 3 * kHeadPatchStride) * 4;
  if (Cout % 8 != 0 || smem > 48 * 1024) return DFIR_ERR_ARG;
  const long long nitems = static_cast<long long>(B) * H * ((W + 127) / 128);
  if (nitems == 0) return DFIR_OK;
  if (nitems > 0x7fffffffll) return DFIR_ERR_ARG;
  const long long nblocks = std::min<long long>(nitems, 148 * 4);
  head_conv_kernel<<<static_cast<unsigned>(nblocks), 256, smem, s>>>(x, wp, bias, out_f32, out_bf16, B, Cin, H, W, Cout,
                                                                      out_lo, lo8);
  return ok_or_cuda();
}

int conv3x3_f32(const float* in, const float* wp, const float* bias, const float* skip, float* out, int B, int H,
                int W, int Cin, int Cout, int relu, int ps_r, int out_nchw, cudaStream_t s, const float* mask) {
  if (Cin % 4 != 0) return DFIR_ERR_ARG;
  if (ps_r > 1 && (Cout % (ps_r * ps_r) != 0 || skip != nullptr || out_nchw)) return DFIR_ERR_ARG;
  const int n_oct = (Cout + 7) / 8;
  const long long nthreads = static_cast<long long>(B) * H * ((W + 3) / 4) * n_oct;
  if (nthreads == 0) return DFIR_OK;
  conv3x3_f32_kernel<<<static_cast<unsigned>((nthreads + 255) / 256), 256, 0, s>>>(in, wp, bias, skip, out, B, H, W,
                                                                                  Cin, Cout, relu, ps_r, out_nchw, mask);
  return ok_or_cuda();
}

int postprocess_rgb(const float* x, float* rgb, float* ycc, int B, long long HW, float lo, float hi, cudaStream_t s) {
  if (B <= 0 || HW <= 0) return DFIR_OK;
  dim3 grid(static_cast<unsigned>(std::min<long long>((HW + 255) / 256, 1184 / std::max(1, std::min(B, 8)) + 1)), B);
  postprocess_rgb_kernel<<<grid, 256, 0, s>>>(x, rgb, ycc, HW, lo, hi, static_cast<float>(128. * (1. / 255)));
  return ok_or_cuda();
}

int pool_rows_f32(const float* in, float* pool_rows, int B, int H, int W, int C, cudaStream_t s) {
  if (B * H == 0) return DFIR_OK;
  pool_rows_f32_kernel<<<B * H, C <= 256 ? C : 256, 0, s>>>(in, pool_rows, W, C);
  return ok_or_cuda();
}

int meta_attention(const float* meta, const float* w1, const float* b1, const float* w2, const float* b2, float* out,
                   int nblk, int B, int M, int Hid, int C, int relu, const int* blk_enabled, float out_scale,
                   cudaStream_t s) {
  if (nblk == 0 || B == 0) return DFIR_OK;
  if ((M + Hid) * 4 > 48 * 1024) return DFIR_ERR_ARG;
  dim3 grid(nblk, B);
  meta_attention_kernel<<<grid, C <= 256 ? C : 256, (M + Hid) * 4, s>>>(meta, w1, b1, w2, b2, out, B, M, Hid, C, relu,
                                                                       blk_enabled, out_scale);
  return ok_or_cuda();
}

int ca_from_stats(const float* pool_rows, const float* col_first, const float* col_last, const void* w2_packed,
                  const float* bias2, const AttnParams& ap, const float* attributes, const float* sq, float* svec, int B,
                  int H, int W, cudaStream_t s) {
  if (B == 0) return DFIR_OK;
  if (ap.C != 64 || ap.A > 512 || ap.M > 448 || ap.style == DFIR_STYLE_NONE) return DFIR_ERR_ARG;
  ca_from_stats_kernel<<<B, 256, 0, s>>>(pool_rows, col_first, col_last, reinterpret_cast<const __nv_bfloat16*>(w2_packed),
                                         bias2, ap, attributes, sq, svec, H, W, (W + 127) / 128);
  return ok_or_cuda();
}

int scale_residual(const void* r, int r_is_bf16, const float* x_in, const float* pool_rows, int pool_nrows,
                   const AttnParams& ap, const float* attributes, const float* sq, float res_scale, float* x_out,
                   __nv_bfloat16* x_out_bf16, int B, int H, int W, int C, cudaStream_t s, float* y_out, const float* pa) {
  if (B == 0 || H * W == 0) return DFIR_OK;
  if (pa != nullptr && (C != 64 || ap.style == DFIR_STYLE_NONE)) return DFIR_ERR_ARG;
  if (C % 8 != 0 || C > 256 || 256 % C != 0 || ap.A > 512 || ap.M > 448) return DFIR_ERR_ARG;
  const long long nvec = static_cast<long long>(H) * W * (C / 8);
  long long per_img = (nvec + 256 * 8 - 1) / (256 * 8);
  const long long cap = std::max(1, 296 / std::max(1, B));  // 128 registers -> 2 CTAs per SM: stay within ONE wave of 296
  if (per_img > cap) per_img = cap;
  if (per_img < 1) per_img = 1;
  dim3 grid(static_cast<unsigned>(per_img), B);
  const int HW = H * W;
  if (r_is_bf16)
    return launch_pdl(PDL_SIMT, scale_residual_kernel<true>, grid, dim3(256), 0, s, r, x_in, pool_rows, pool_nrows, ap, attributes, sq,
                      res_scale, x_out, x_out_bf16, HW, C, y_out, pa) == cudaSuccess ? DFIR_OK : DFIR_ERR_CUDA;
  return launch_pdl(PDL_SIMT, scale_residual_kernel<false>, grid, dim3(256), 0, s, r, x_in, pool_rows, pool_nrows, ap, attributes, sq,
                    res_scale, x_out, x_out_bf16, HW, C, y_out, pa) == cudaSuccess ? DFIR_OK : DFIR_ERR_CUDA;
}

}  // namespace dfir
