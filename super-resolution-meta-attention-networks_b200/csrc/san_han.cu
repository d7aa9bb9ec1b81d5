// Kernels of the Q-HAN and Q-SAN specific layers (all HBM / latency bound, fp32, NHWC):
//   LAM  layer attention           (reference advanced/HAN_blocks.py:7-37)
//   CSAM channel-spatial attention (reference advanced/HAN_blocks.py:40-76)
//   Covpool + Newton-Schulz sqrt + SOCA MLP (reference advanced/mpncov.py:12-76, advanced/SAN_blocks.py:244-302)
//   region non-local attention     (reference advanced/SAN_blocks.py:11-148, 305-336)
// Every reduction uses a fixed two-stage order, so results are bit-reproducible and independent of the batch.
#include "kernels.h"
#include "attn.cuh"

#include <algorithm>

namespace dfir {

namespace {

inline int ok_or_cuda2() { return cudaGetLastError() == cudaSuccess ? DFIR_OK : DFIR_ERR_CUDA; }

// ------------------------------------------------------------------------------------------------
// small elementwise helpers
// ------------------------------------------------------------------------------------------------
__global__ void f32_to_bf16_kernel(const float4* __restrict__ in, uint2* __restrict__ out, long long n4) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 v = in[i];
    __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
    uint2 o;
    o.x = *reinterpret_cast<uint32_t*>(&a);
    o.y = *reinterpret_cast<uint32_t*>(&b);
    out[i] = o;
  }
}

// fp32 <-> the hi / 8-bit lo stream format (kernels.h, EPI_SCALE_SKIP_HL8): bits(X) = (hi << 16) + (q << 8), X = x rounded to 24
// bits; 64-channel pixels; the lo plane stores a pixel's bytes in accumulator-fragment order: byte cq * 16 + 2 n + e holds
// channel 8 n + 2 cq + e.  One thread = 4 consecutive channels 4 m .. 4 m + 3 of a pixel (n = m >> 1, cq = 2 (m & 1) + {0, 1}).
__global__ void stream_encode_hl8_kernel(const float4* __restrict__ in, uint2* __restrict__ hi, unsigned char* __restrict__ lo,
                                         long long n4) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 v = in[i];
    const float f[4] = {v.x, v.y, v.z, v.w};
    uint32_t hb[4], q[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t t = __float_as_uint(f[k]) + 0x80u;
      hb[k] = (t + 0x8000u) & 0xffff0000u;
      q[k] = ((t - hb[k]) >> 8) & 0xffu;
    }
    hi[i] = make_uint2((hb[0] >> 16) | hb[1], (hb[2] >> 16) | hb[3]);
    const int m = static_cast<int>(i & 15);
    unsigned char* px = lo + (i >> 4) * 64 + 2 * (m >> 1) + 32 * (m & 1);
    *reinterpret_cast<unsigned short*>(px) = static_cast<unsigned short>(q[0] | (q[1] << 8));
    *reinterpret_cast<unsigned short*>(px + 16) = static_cast<unsigned short>(q[2] | (q[3] << 8));
  }
}

__global__ void stream_decode_hl8_kernel(const uint2* __restrict__ hi, const unsigned char* __restrict__ lo,
                                         float4* __restrict__ out, long long n4) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const uint2 h = hi[i];
    const int m = static_cast<int>(i & 15);
    const unsigned char* px = lo + (i >> 4) * 64 + 2 * (m >> 1) + 32 * (m & 1);
    const uint32_t qa = *reinterpret_cast<const unsigned short*>(px), qb = *reinterpret_cast<const unsigned short*>(px + 16);
    const uint32_t q[4] = {qa & 0xffu, qa >> 8, qb & 0xffu, qb >> 8};
    const uint32_t hb[4] = {h.x << 16, h.x & 0xffff0000u, h.y << 16, h.y & 0xffff0000u};
    float f[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int qs = static_cast<int>(static_cast<int8_t>(q[k]));
      f[k] = __uint_as_float(hb[k] + static_cast<uint32_t>(qs << 8));
    }
    out[i] = make_float4(f[0], f[1], f[2], f[3]);
  }
}

// out = x * s[b][c] (+ add), per-image channel scale (SOCA `y * x`, SAN_blocks.py:302)
__global__ void channel_scale_kernel(const float4* __restrict__ x, const float* __restrict__ s,
                                     const float4* __restrict__ add, float alpha, float4* __restrict__ out,
                                     long long per_img4, int C4, long long n4) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int b = static_cast<int>(i / per_img4);
    const int c4 = static_cast<int>(i % C4);
    float4 v = x[i];
    if (s != nullptr) {
      const float4 sv = *reinterpret_cast<const float4*>(s + static_cast<size_t>(b) * C4 * 4 + c4 * 4);
      v.x *= sv.x; v.y *= sv.y; v.z *= sv.z; v.w *= sv.w;
    }
    if (add != nullptr) {
      const float4 a = add[i];
      v.x = fmaf(alpha, a.x, v.x); v.y = fmaf(alpha, a.y, v.y); v.z = fmaf(alpha, a.z, v.z); v.w = fmaf(alpha, a.w, v.w);
    }
    out[i] = v;
  }
}

// ------------------------------------------------------------------------------------------------
// LAM: energy[b][i][j] = sum over (pixel, channel) of X_i * X_j ; attention = softmax_j(max_j E[i] - E[i][j]);
//      out[b][p][i*C + c] = gamma * sum_j att[i][j] X_j[b][p][c] + X_i[b][p][c]
// stack: N feature maps, map n at stack + n * map_stride (each [B][HW][C] fp32).
// ------------------------------------------------------------------------------------------------
constexpr int kLamMaxN = 16;

__global__ void __launch_bounds__(256)
lam_gram_kernel(const float* __restrict__ stack, long long map_stride, float* __restrict__ partial, int N, int HWC,
                int nchunk) {
  __shared__ float red[8][kLamMaxN * kLamMaxN];
  const int b = blockIdx.y, chunk = blockIdx.x;
  const long long per = (static_cast<long long>(HWC) + nchunk - 1) / nchunk;
  const long long e0 = chunk * per, e1 = min(static_cast<long long>(HWC), e0 + per);
  float acc[kLamMaxN * (kLamMaxN + 1) / 2];
#pragma unroll
  for (int k = 0; k < kLamMaxN * (kLamMaxN + 1) / 2; ++k) acc[k] = 0.f;
  const float* base = stack + static_cast<size_t>(b) * HWC;
  for (long long e = e0 + threadIdx.x; e < e1; e += 256) {
    float v[kLamMaxN];
#pragma unroll
    for (int n = 0; n < kLamMaxN; ++n) v[n] = n < N ? base[n * map_stride + e] : 0.f;
    int k = 0;
#pragma unroll
    for (int i = 0; i < kLamMaxN; ++i)
#pragma unroll
      for (int j = i; j < kLamMaxN; ++j, ++k) acc[k] = fmaf(v[i], v[j], acc[k]);
  }
  // block reduction in a fixed order: warp shuffle tree, then 8 warps summed sequentially
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int k = 0;
  for (int i = 0; i < kLamMaxN; ++i)
    for (int j = i; j < kLamMaxN; ++j, ++k) {
      float v = acc[k];
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) red[warp][i * kLamMaxN + j] = v;
    }
  __syncthreads();
  for (int idx = threadIdx.x; idx < N * N; idx += 256) {
    const int i = idx / N, j = idx % N;
    const int a = i <= j ? i : j, c = i <= j ? j : i;
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w][a * kLamMaxN + c];
    partial[(static_cast<size_t>(b) * nchunk + chunk) * N * N + idx] = t;
  }
}

__global__ void lam_attention_kernel(const float* __restrict__ partial, float* __restrict__ att, int N, int nchunk) {
  __shared__ float E[kLamMaxN * kLamMaxN];
  const int b = blockIdx.x;
  for (int idx = threadIdx.x; idx < N * N; idx += blockDim.x) {
    float t = 0.f;
    for (int c = 0; c < nchunk; ++c) t += partial[(static_cast<size_t>(b) * nchunk + c) * N * N + idx];
    E[idx] = t;
  }
  __syncthreads();
  if (threadIdx.x < N) {
    const int i = threadIdx.x;
    float mx = -3.4e38f;
    for (int j = 0; j < N; ++j) mx = fmaxf(mx, E[i * N + j]);
    // energy_new = max - E ; softmax over j (stabilised by its own maximum, like torch.softmax)
    float m2 = -3.4e38f;
    for (int j = 0; j < N; ++j) m2 = fmaxf(m2, mx - E[i * N + j]);
    float sum = 0.f;
    for (int j = 0; j < N; ++j) sum += expf((mx - E[i * N + j]) - m2);
    for (int j = 0; j < N; ++j) att[(static_cast<size_t>(b) * N + i) * N + j] = expf((mx - E[i * N + j]) - m2) / sum;
  }
}

__global__ void __launch_bounds__(256)
lam_apply_kernel(const float* __restrict__ stack, long long map_stride, const float* __restrict__ att, float gamma,
                 float* __restrict__ out, int N, int HW, int C) {
  __shared__ float a_s[kLamMaxN * kLamMaxN];
  const int b = blockIdx.y;
  for (int i = threadIdx.x; i < N * N; i += 256) a_s[i] = att[static_cast<size_t>(b) * N * N + i];
  __syncthreads();
  const long long per_img = static_cast<long long>(HW) * C;
  for (long long e = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; e < per_img;
       e += static_cast<long long>(gridDim.x) * 256) {
    const long long p = e / C;
    const int c = static_cast<int>(e % C);
    float v[kLamMaxN];
    for (int n = 0; n < N; ++n) v[n] = stack[n * map_stride + static_cast<size_t>(b) * per_img + e];
    for (int i = 0; i < N; ++i) {
      float t = 0.f;
      for (int j = 0; j < N; ++j) t = fmaf(a_s[i * N + j], v[j], t);
      out[(static_cast<size_t>(b) * HW + p) * (static_cast<size_t>(N) * C) + i * C + c] = fmaf(gamma, t, v[i]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// CSAM: out = x * (gamma * sigmoid(conv3d_3x3x3(x viewed as a (C,H,W) volume))) + x
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
csam_kernel(const float* __restrict__ x, const float* __restrict__ w27, float bias, float gamma,
            float* __restrict__ out, int B, int H, int W, int C) {
  __shared__ float ws[27];
  if (threadIdx.x < 27) ws[threadIdx.x] = w27[threadIdx.x];
  __syncthreads();
  const long long n = static_cast<long long>(B) * H * W * C;
  for (long long e = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; e < n;
       e += static_cast<long long>(gridDim.x) * 256) {
    const int c = static_cast<int>(e % C);
    const long long pix = e / C;
    const int xw = static_cast<int>(pix % W);
    const int y = static_cast<int>((pix / W) % H);
    const int b = static_cast<int>(pix / (static_cast<long long>(W) * H));
    float acc = bias;
#pragma unroll
    for (int dc = 0; dc < 3; ++dc) {
      const int cc = c + dc - 1;
      if (cc < 0 || cc >= C) continue;
#pragma unroll
      for (int dy = 0; dy < 3; ++dy) {
        const int yy = y + dy - 1;
        if (yy < 0 || yy >= H) continue;
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          const int xx = xw + dx - 1;
          if (xx < 0 || xx >= W) continue;
          acc = fmaf(ws[(dc * 3 + dy) * 3 + dx], x[((static_cast<size_t>(b) * H + yy) * W + xx) * C + cc], acc);
        }
      }
    }
    const float g = gamma / (1.f + expf(-acc));
    const float v = x[e];
    out[e] = fmaf(v, g, v);
  }
}

// ------------------------------------------------------------------------------------------------
// Covpool: Sigma = X X^T / M - (X1)(X1)^T / M^2 over the (optionally centre-cropped) window, C = 64.
// stage 1: each CTA accumulates a 64x64 Gram + channel sums over a chunk of pixels; stage 2: fixed-order sum.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
covpool_partial_kernel(const float* __restrict__ x, float* __restrict__ partial, int H, int W, int y0, int x0, int h1,
                       int w1, int nchunk) {
  __shared__ float tile[64][65];  // [pixel][channel], padded
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int M = h1 * w1;
  const int per = (M + nchunk - 1) / nchunk;
  const int p0 = chunk * per, p1 = min(M, p0 + per);
  const int ti = (threadIdx.x >> 4) * 4, tj = (threadIdx.x & 15) * 4;  // 4x4 output tile of this thread
  float acc[4][4] = {};
  float csum = 0.f;  // thread c < 64: sum of channel c
  for (int base = p0; base < p1; base += 64) {
    const int np = min(64, p1 - base);
    for (int idx = threadIdx.x; idx < 64 * 64; idx += 256) {
      const int pp = idx >> 6, c = idx & 63;
      float v = 0.f;
      if (pp < np) {
        const int q = base + pp;
        const int yy = y0 + q / w1, xx = x0 + q % w1;
        v = x[((static_cast<size_t>(b) * H + yy) * W + xx) * 64 + c];
      }
      tile[pp][c] = v;
    }
    __syncthreads();
    for (int pp = 0; pp < np; ++pp) {
      float a[4], bb[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) { a[k] = tile[pp][ti + k]; bb[k] = tile[pp][tj + k]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    if (threadIdx.x < 64)
      for (int pp = 0; pp < np; ++pp) csum += tile[pp][threadIdx.x];
    __syncthreads();
  }
  float* o = partial + (static_cast<size_t>(b) * nchunk + chunk) * (64 * 64 + 64);
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) o[(ti + i) * 64 + tj + j] = acc[i][j];
  if (threadIdx.x < 64) o[64 * 64 + threadIdx.x] = csum;
}

__global__ void covpool_final_kernel(const float* __restrict__ partial, float* __restrict__ cov, int M, int nchunk) {
  __shared__ float s[64];
  const int b = blockIdx.x;  // blockIdx.y: slice of 256 of the 4096 matrix entries (fixed summation order per entry)
  if (threadIdx.x < 64) {
    float t = 0.f;
    for (int c = 0; c < nchunk; ++c) t += partial[(static_cast<size_t>(b) * nchunk + c) * (64 * 64 + 64) + 64 * 64 + threadIdx.x];
    s[threadIdx.x] = t;
  }
  __syncthreads();
  const float inv = 1.f / static_cast<float>(M);
  for (int idx = blockIdx.y * 256 + threadIdx.x; idx < 64 * 64; idx += gridDim.y * 256) {
    float t = 0.f;
    for (int c = 0; c < nchunk; ++c) t += partial[(static_cast<size_t>(b) * nchunk + c) * (64 * 64 + 64) + idx];
    cov[static_cast<size_t>(b) * 4096 + idx] = t * inv - (s[idx >> 6] * inv) * (s[idx & 63] * inv);
  }
}

// ------------------------------------------------------------------------------------------------
// Newton-Schulz matrix square root (5 iterations) + column mean + SOCA MLP, one CTA (256 threads) per image.
//   A = Sigma / tr ; ZY = (3I - A)/2 ; Y = A ZY ; Z = ZY ; 3x { ZY = (3I - Z Y)/2 ; Y = Y ZY ; Z = ZY Z } ;
//   S = Y (3I - Z Y) / 2 * sqrt(tr) ; v = mean over dim 1 of S ; s = sigmoid(W2 relu(W1 v))
// ------------------------------------------------------------------------------------------------
__device__ void mm64(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ Cc, float alpha,
                     float diag) {
  // C = alpha * (A @ B) + diag * I   (row-major 64x64 in shared memory; each thread a 4x4 tile)
  const int ti = (threadIdx.x >> 4) * 4, tj = (threadIdx.x & 15) * 4;
  float acc[4][4] = {};
  for (int k = 0; k < 64; ++k) {
    float a[4], b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { a[i] = A[(ti + i) * 64 + k]; b[i] = B[k * 64 + tj + i]; }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
  }
  __syncthreads();  // all reads of A/B done before C (which may alias neither, but callers rotate buffers) is written
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) Cc[(ti + i) * 64 + tj + j] = alpha * acc[i][j] + ((ti + i) == (tj + j) ? diag : 0.f);
  __syncthreads();
}

__global__ void __launch_bounds__(256)
soca_kernel(const float* __restrict__ cov, const float* __restrict__ mlp, int R, float* __restrict__ svec, int iters) {
  extern __shared__ float sm[];
  float* Y = sm;             // 64x64
  float* Z = sm + 4096;
  float* T = sm + 8192;      // scratch
  float* U = sm + 12288;     // scratch
  __shared__ float y_s[64], s_s[64], tmp[256], trace;
  const int b = blockIdx.x;
  const float* S = cov + static_cast<size_t>(b) * 4096;
  for (int i = threadIdx.x; i < 4096; i += 256) T[i] = S[i];
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 64; ++i) t += T[i * 65];
    trace = t;
  }
  __syncthreads();
  const float inv = 1.f / trace;
  // U = A = Sigma / tr ; Z = ZY = 0.5 (3I - A)
  for (int i = threadIdx.x; i < 4096; i += 256) {
    const float a = T[i] * inv;
    U[i] = a;
    Z[i] = 0.5f * (((i >> 6) == (i & 63) ? 3.f : 0.f) - a);
  }
  __syncthreads();
  mm64(U, Z, Y, 1.f, 0.f);  // Y0 = A ZY
  for (int it = 1; it < iters - 1; ++it) {
    mm64(Z, Y, T, -0.5f, 1.5f);  // ZY = 0.5 (3I - Z Y)
    mm64(Y, T, U, 1.f, 0.f);     // Y' = Y ZY
    mm64(T, Z, Y, 1.f, 0.f);     // Z' = ZY Z   (into Y's buffer, then swap roles)
    float* t0 = Y; Y = U; U = Z; Z = t0;  // Y <- Y', Z <- Z'
  }
  mm64(Z, Y, T, -0.5f, 1.5f);    // 0.5 (3I - Z Y)
  mm64(Y, T, U, 1.f, 0.f);       // Y (..)/2 ... U = final / sqrt(tr)
  if (threadIdx.x < 64) {
    float t = 0.f;
    for (int i = 0; i < 64; ++i) t += U[i * 64 + threadIdx.x];  // torch.mean(S, 1): mean over the row index
    y_s[threadIdx.x] = t * (1.f / 64.f) * sqrtf(trace);
  }
  __syncthreads();
  // FC(64 -> R) ReLU FC(R -> 64) sigmoid, parameter order of DFIR_STYLE_STANDARD
  attn_vector(BlockGroup{}, DFIR_STYLE_STANDARD, mlp, 64, R, 0, nullptr, y_s, s_s, tmp);
  if (threadIdx.x < 64) svec[static_cast<size_t>(b) * 64 + threadIdx.x] = s_s[threadIdx.x];
}

// ------------------------------------------------------------------------------------------------
// Covpool.backward (advanced/mpncov.py:35-47): grad_x = (G + G^T) X I_hat, I_hat = I/M - 11^T/M^2, i.e.
//   grad_x[p][c] = (1/M) sum_d (G[c][d] + G[d][c]) (x[p][d] - mean_p x[.][d])
// without the MxM matrix.  Stage 1: per-chunk channel sums (fixed order); stage 2: one 64x64 by 64xM product, streamed.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
channel_sum_partial_kernel(const float* __restrict__ x, float* __restrict__ partial, int H, int W, int y0, int x0, int h1,
                           int w1, int nchunk) {
  __shared__ float red[4][64];
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int M = h1 * w1;
  const int per = (M + nchunk - 1) / nchunk;
  const int p0 = chunk * per, p1 = min(M, p0 + per);
  const int c = threadIdx.x & 63, lane4 = threadIdx.x >> 6;
  float t = 0.f;
  for (int q = p0 + lane4; q < p1; q += 4) {
    const int yy = y0 + q / w1, xx = x0 + q % w1;
    t += x[((static_cast<size_t>(b) * H + yy) * W + xx) * 64 + c];
  }
  red[lane4][c] = t;
  __syncthreads();
  if (threadIdx.x < 64)
    partial[(static_cast<size_t>(b) * nchunk + chunk) * 64 + c] = (red[0][c] + red[1][c]) + (red[2][c] + red[3][c]);
}

__global__ void __launch_bounds__(256)
covpool_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gcov, const float* __restrict__ partial,
                   float* __restrict__ gx, int H, int W, int y0, int x0, int h1, int w1, int nchunk, int nsum) {
  __shared__ float S[64][65];     // (G + G^T) / M, [d][c] (symmetric)
  __shared__ float tile[64][65];  // [pixel][channel] of x - mean
  __shared__ float mean[64];
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int M = h1 * w1;
  const float inv = 1.f / static_cast<float>(M);
  if (threadIdx.x < 64) {
    float t = 0.f;
    for (int k = 0; k < nsum; ++k) t += partial[(static_cast<size_t>(b) * nsum + k) * 64 + threadIdx.x];
    mean[threadIdx.x] = t * inv;
  }
  const float* g = gcov + static_cast<size_t>(b) * 4096;
  for (int idx = threadIdx.x; idx < 4096; idx += 256) {
    const int i = idx >> 6, j = idx & 63;
    S[i][j] = (g[i * 64 + j] + g[j * 64 + i]) * inv;
  }
  __syncthreads();
  const int per = (M + nchunk - 1) / nchunk;
  const int p0 = chunk * per, p1 = min(M, p0 + per);
  const int ti = (threadIdx.x >> 4) * 4, tj = (threadIdx.x & 15) * 4;  // 4 pixels x 4 output channels per thread
  for (int base = p0; base < p1; base += 64) {
    const int np = min(64, p1 - base);
    for (int idx = threadIdx.x; idx < 64 * 64; idx += 256) {
      const int pp = idx >> 6, c = idx & 63;
      float v = 0.f;
      if (pp < np) {
        const int q = base + pp;
        const int yy = y0 + q / w1, xx = x0 + q % w1;
        v = x[((static_cast<size_t>(b) * H + yy) * W + xx) * 64 + c] - mean[c];
      }
      tile[pp][c] = v;
    }
    __syncthreads();
    float acc[4][4] = {};
    for (int d = 0; d < 64; ++d) {
      float a[4], w[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) { a[k] = tile[ti + k][d]; w[k] = S[d][tj + k]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int pp = ti + i;
      if (pp < np) {
        const int q = base + pp;
        const int yy = y0 + q / w1, xx = x0 + q % w1;
        *reinterpret_cast<float4*>(gx + ((static_cast<size_t>(b) * H + yy) * W + xx) * 64 + tj) =
            make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// Sqrtm (Newton-Schulz, advanced/mpncov.py:49-112) as stand-alone operators: forward returning the matrix, and the
// reference's hand-derived backward.  One CTA (256 threads) per image, all 64x64 matrices in shared memory; the forward
// iterates Y_i, Z_i are kept in a small global stash (written and read by the same CTA, through L2).
// ------------------------------------------------------------------------------------------------
__device__ void mm64x(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ Cc, float alpha,
                      float beta, float diag) {
  // C = alpha * (A @ B) + beta * C + diag * I   (row-major 64x64 in shared memory; C aliases neither A nor B)
  const int ti = (threadIdx.x >> 4) * 4, tj = (threadIdx.x & 15) * 4;
  float acc[4][4] = {};
  for (int k = 0; k < 64; ++k) {
    float a[4], b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { a[i] = A[(ti + i) * 64 + k]; b[i] = B[k * 64 + tj + i]; }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int e = (ti + i) * 64 + tj + j;
      const float old = beta != 0.f ? beta * Cc[e] : 0.f;
      Cc[e] = alpha * acc[i][j] + old + ((ti + i) == (tj + j) ? diag : 0.f);
    }
  __syncthreads();
}

__device__ float block_sum256(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  for (int w = 0; w < 8; ++w) t += red[w];
  __syncthreads();
  return t;
}

// forward iterates: on exit Y = Y_{n-2}, Z = Z_{n-2} (n = iters), M2 = 0.5 Y (3I - Z Y) (the result before * sqrt(tr));
// stash (optional): Y_0, Z_0, ..., Y_{n-2}, Z_{n-2} as [2 (n-1)][4096]
__device__ void sqrtm_forward_smem(const float* __restrict__ cov_g, float* A, float*& Y, float*& Z, float* M1, float*& M2,
                                   float*& M3, float* red, float& trace, int iters, float* __restrict__ stash) {
  for (int i = threadIdx.x; i < 4096; i += 256) A[i] = cov_g[i];
  __syncthreads();
  float d = threadIdx.x < 64 ? A[threadIdx.x * 65] : 0.f;
  trace = block_sum256(d, red);
  const float inv = 1.f / trace;
  for (int i = threadIdx.x; i < 4096; i += 256) {
    const float a = A[i] * inv;
    A[i] = a;
    Z[i] = 0.5f * (((i >> 6) == (i & 63) ? 3.f : 0.f) - a);
  }
  __syncthreads();
  mm64x(A, Z, Y, 1.f, 0.f, 0.f);  // Y0 = A ZY
  auto put = [&](int slot, const float* src) {
    if (stash != nullptr)
      for (int i = threadIdx.x; i < 4096; i += 256) stash[static_cast<size_t>(slot) * 4096 + i] = src[i];
  };
  put(0, Y); put(1, Z);
  for (int it = 1; it < iters - 1; ++it) {
    mm64x(Z, Y, M1, -0.5f, 0.f, 1.5f);  // ZY = 0.5 (3I - Z Y)
    mm64x(Y, M1, M2, 1.f, 0.f, 0.f);    // Y' = Y ZY
    mm64x(M1, Z, M3, 1.f, 0.f, 0.f);    // Z' = ZY Z
    float* oy = Y; float* oz = Z;
    Y = M2; Z = M3; M2 = oy; M3 = oz;
    put(2 * it, Y); put(2 * it + 1, Z);
  }
  mm64x(Z, Y, M1, -1.f, 0.f, 3.f);      // 3I - Z Y
  mm64x(Y, M1, M2, 0.5f, 0.f, 0.f);     // 0.5 Y (3I - Z Y)
}

__global__ void __launch_bounds__(256) sqrtm_fwd_kernel(const float* __restrict__ cov, float* __restrict__ out, int iters) {
  extern __shared__ float sm[];
  float *A = sm, *Y = sm + 4096, *Z = sm + 8192, *M1 = sm + 12288, *M2 = sm + 16384, *M3 = sm + 20480;
  __shared__ float red[8];
  float trace;
  const int b = blockIdx.x;
  sqrtm_forward_smem(cov + static_cast<size_t>(b) * 4096, A, Y, Z, M1, M2, M3, red, trace, iters, nullptr);
  const float sq = sqrtf(trace);
  for (int i = threadIdx.x; i < 4096; i += 256) out[static_cast<size_t>(b) * 4096 + i] = M2[i] * sq;
}

__global__ void __launch_bounds__(256)
sqrtm_bwd_kernel(const float* __restrict__ cov, const float* __restrict__ gout, float* __restrict__ gin,
                 float* __restrict__ stash_all, int iters) {
  extern __shared__ float sm[];
  float* A = sm;               // x / trace
  float* Yi = sm + 1 * 4096;   // forward iterate Y_i (rotates with the scratch slots during the recompute)
  float* Zi = sm + 2 * 4096;
  float* M1 = sm + 3 * 4096;
  float* M2 = sm + 4 * 4096;
  float* M3 = sm + 5 * 4096;
  float* dY = sm + 6 * 4096;
  float* dZ = sm + 7 * 4096;
  float* nY = sm + 8 * 4096;
  float* nZ = sm + 9 * 4096;
  __shared__ float red[8];
  const int b = blockIdx.x;
  const float* xg = cov + static_cast<size_t>(b) * 4096;
  const float* gg = gout + static_cast<size_t>(b) * 4096;
  float* stash = stash_all + static_cast<size_t>(b) * 2 * (iters - 1) * 4096;
  float trace;
  sqrtm_forward_smem(xg, A, Yi, Zi, M1, M2, M3, red, trace, iters, stash);
  // now: Yi = Y_{n-2}, Zi = Z_{n-2}, M2 = ZY_final (= y / sqrt(trace)); M1, M3 free
  const float sq = sqrtf(trace);
  float part = 0.f;
  for (int i = threadIdx.x; i < 4096; i += 256) {
    const float g = gg[i];
    part += g * M2[i];
    nY[i] = g * sq;  // der_postCom
  }
  const float aux = block_sum256(part, red) / (2.f * sq);  // der_postComAux
  float* dpc = nY;
  // dldY = 0.5 (dpc (3I - Y Z) - Z Y dpc) ; dldZ = -0.5 Y dpc Y
  mm64x(Yi, Zi, M1, -1.f, 0.f, 3.f);   // YZ = 3I - Y Z
  mm64x(dpc, M1, dY, 0.5f, 0.f, 0.f);
  mm64x(Zi, Yi, M3, 1.f, 0.f, 0.f);    // ZY
  mm64x(M3, dpc, dY, -0.5f, 1.f, 0.f);
  mm64x(Yi, dpc, M3, 1.f, 0.f, 0.f);
  mm64x(M3, Yi, dZ, -0.5f, 0.f, 0.f);
  for (int i = iters - 3; i >= 0; --i) {
    for (int k = threadIdx.x; k < 4096; k += 256) {  // Y_i, Z_i back from the stash (through L2: written by this CTA)
      Yi[k] = __ldcg(stash + static_cast<size_t>(2 * i) * 4096 + k);
      Zi[k] = __ldcg(stash + static_cast<size_t>(2 * i + 1) * 4096 + k);
    }
    __syncthreads();
    mm64x(Yi, Zi, M1, -1.f, 0.f, 3.f);  // YZ = 3I - Y_i Z_i
    mm64x(Zi, Yi, M2, 1.f, 0.f, 0.f);   // ZY = Z_i Y_i
    // dldY_ = 0.5 (dldY YZ - Z_i dldZ Z_i - ZY dldY)
    mm64x(dY, M1, nY, 0.5f, 0.f, 0.f);
    mm64x(Zi, dZ, M3, 1.f, 0.f, 0.f);
    mm64x(M3, Zi, nY, -0.5f, 1.f, 0.f);
    mm64x(M2, dY, nY, -0.5f, 1.f, 0.f);
    // dldZ_ = 0.5 (YZ dldZ - Y_i dldY Y_i - dldZ ZY)
    mm64x(M1, dZ, nZ, 0.5f, 0.f, 0.f);
    mm64x(Yi, dY, M3, 1.f, 0.f, 0.f);
    mm64x(M3, Yi, nZ, -0.5f, 1.f, 0.f);
    mm64x(dZ, M2, nZ, -0.5f, 1.f, 0.f);
    float* t = dY; dY = nY; nY = t;
    t = dZ; dZ = nZ; nZ = t;
  }
  // der_NSiter = 0.5 (dldY (3I - A) - dldZ - A dldY)
  for (int k = threadIdx.x; k < 4096; k += 256) M1[k] = ((k >> 6) == (k & 63) ? 3.f : 0.f) - A[k];
  __syncthreads();
  mm64x(dY, M1, nY, 0.5f, 0.f, 0.f);
  mm64x(A, dY, nY, -0.5f, 1.f, 0.f);
  float gpart = 0.f;
  for (int k = threadIdx.x; k < 4096; k += 256) {
    const float d = nY[k] - 0.5f * dZ[k];
    nY[k] = d;
    gpart += d * xg[k];
  }
  const float grad_aux = block_sum256(gpart, red);
  const float diag = aux - grad_aux / (trace * trace);
  for (int k = threadIdx.x; k < 4096; k += 256)
    gin[static_cast<size_t>(b) * 4096 + k] = nY[k] / trace + ((k >> 6) == (k & 63) ? diag : 0.f);
}

// ------------------------------------------------------------------------------------------------
// Region non-local attention (embedded Gaussian, inter channels D = 8, always 2x2 max-pooled keys/values).
// nl_project: theta / phi / g 1x1 convs (64 -> 8 each) for every pixel: out [B][H][W][24]
// nl_pool   : 2x2 max-pool of phi and g inside each of the 4 regions: keys [B][4][Nk_max][16]
// nl_attend : one thread per query pixel, keys of its region streamed through shared memory, online softmax;
//             z = W y + b + x
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
nl_project_kernel(const float* __restrict__ x, const float* __restrict__ wq, const float* __restrict__ bq,
                  float* __restrict__ proj, long long npix) {
  __shared__ float ws[24 * 64 + 24];
  for (int i = threadIdx.x; i < 24 * 64; i += 256) ws[i] = wq[i];
  if (threadIdx.x < 24) ws[24 * 64 + threadIdx.x] = bq[threadIdx.x];
  __syncthreads();
  for (long long p = static_cast<long long>(blockIdx.x) * 256 + threadIdx.x; p < npix;
       p += static_cast<long long>(gridDim.x) * 256) {
    float v[64];
    const float4* px = reinterpret_cast<const float4*>(x + p * 64);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float4 t = px[i];
      v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
    }
    for (int o = 0; o < 24; ++o) {
      float acc = ws[24 * 64 + o];
#pragma unroll
      for (int c = 0; c < 64; ++c) acc = fmaf(ws[o * 64 + c], v[c], acc);
      proj[p * 24 + o] = acc;
    }
  }
}

struct Region { int y0, y1, x0, x1; };
__device__ __forceinline__ Region region_of(int r, int H, int W) {
  const int H1 = H / 2, W1 = W / 2;  // int(H / 2), int(W / 2) in the reference
  Region g;
  g.y0 = (r & 1) ? H1 : 0; g.y1 = (r & 1) ? H : H1;   // regions ordered (top-left, bottom-left, top-right, bottom-right)
  g.x0 = (r & 2) ? W1 : 0; g.x1 = (r & 2) ? W : W1;
  return g;
}

__global__ void nl_pool_kernel(const float* __restrict__ proj, float* __restrict__ keys, int B, int H, int W,
                               int nk_max) {
  const int r = blockIdx.y, b = blockIdx.z;
  const Region g = region_of(r, H, W);
  const int ph = (g.y1 - g.y0) / 2, pw = (g.x1 - g.x0) / 2;
  for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < ph * pw; k += gridDim.x * blockDim.x) {
    const int ky = k / pw, kx = k % pw;
    float m[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) m[i] = -3.4e38f;
    for (int dy = 0; dy < 2; ++dy)
      for (int dx = 0; dx < 2; ++dx) {
        const float* p = proj + ((static_cast<size_t>(b) * H + g.y0 + 2 * ky + dy) * W + g.x0 + 2 * kx + dx) * 24 + 8;
#pragma unroll
        for (int i = 0; i < 16; ++i) m[i] = fmaxf(m[i], p[i]);  // [phi(8) | g(8)]
      }
    float* o = keys + ((static_cast<size_t>(b) * 4 + r) * nk_max + k) * 16;
#pragma unroll
    for (int i = 0; i < 16; ++i) o[i] = m[i];
  }
}

__global__ void __launch_bounds__(128)
nl_attend_kernel(const float* __restrict__ x, const float* __restrict__ proj, const float* __restrict__ keys,
                 const float* __restrict__ wW, const float* __restrict__ bW, float* __restrict__ out, int H, int W,
                 int nk_max) {
  __shared__ float ks[256 * 16];
  __shared__ float w_s[64 * 8 + 64];
  const int r = blockIdx.y, b = blockIdx.z;
  const Region g = region_of(r, H, W);
  const int rh = g.y1 - g.y0, rw = g.x1 - g.x0;
  const int nq = rh * rw, nk = (rh / 2) * (rw / 2);
  for (int i = threadIdx.x; i < 64 * 8; i += 128) w_s[i] = wW[i];
  if (threadIdx.x < 64) w_s[512 + threadIdx.x] = bW[threadIdx.x];
  const int q = blockIdx.x * 128 + threadIdx.x;
  const bool active = q < nq;
  const int qy = active ? g.y0 + q / rw : g.y0, qx = active ? g.x0 + q % rw : g.x0;
  const size_t pix = (static_cast<size_t>(b) * H + qy) * W + qx;
  float th[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) th[i] = proj[pix * 24 + i];
  float mx = -3.4e38f, den = 0.f, y[8] = {};
  const float* kbase = keys + (static_cast<size_t>(b) * 4 + r) * nk_max * 16;
  for (int k0 = 0; k0 < nk; k0 += 256) {
    const int nkt = min(256, nk - k0);
    __syncthreads();
    for (int i = threadIdx.x; i < nkt * 16; i += 128) ks[i] = kbase[static_cast<size_t>(k0) * 16 + i];
    __syncthreads();
    for (int k = 0; k < nkt; ++k) {
      const float* kk = ks + k * 16;
      float sc = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) sc = fmaf(th[i], kk[i], sc);
      if (sc > mx) {  // online softmax: rescale the running sums when the maximum grows
        const float f = expf(mx - sc);
        den *= f;
#pragma unroll
        for (int i = 0; i < 8; ++i) y[i] *= f;
        mx = sc;
      }
      const float e = expf(sc - mx);
      den += e;
#pragma unroll
      for (int i = 0; i < 8; ++i) y[i] = fmaf(e, kk[8 + i], y[i]);
    }
  }
  if (!active) return;
  const float inv = nk > 0 ? 1.f / den : 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) y[i] *= inv;
  const float* px = x + pix * 64;
  float* po = out + pix * 64;
  for (int c = 0; c < 64; ++c) {
    float acc = w_s[512 + c];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc = fmaf(w_s[c * 8 + i], y[i], acc);
    po[c] = acc + px[c];
  }
}

}  // namespace

int f32_to_bf16(const float* in, __nv_bfloat16* out, long long n, cudaStream_t s) {
  if (n % 4 != 0) return DFIR_ERR_ARG;
  if (n == 0) return DFIR_OK;
  const long long n4 = n / 4;
  f32_to_bf16_kernel<<<static_cast<unsigned>(std::min<long long>((n4 + 255) / 256, 148 * 8)), 256, 0, s>>>(
      reinterpret_cast<const float4*>(in), reinterpret_cast<uint2*>(out), n4);
  return ok_or_cuda2();
}

int stream_encode_hl8(const float* in, void* hi, void* lo8, long long n, cudaStream_t s) {
  if (n % 64 != 0) return DFIR_ERR_ARG;  // whole 64-channel pixels
  if (n == 0) return DFIR_OK;
  const long long n4 = n / 4;
  stream_encode_hl8_kernel<<<static_cast<unsigned>(std::min<long long>((n4 + 255) / 256, 148 * 8)), 256, 0, s>>>(
      reinterpret_cast<const float4*>(in), reinterpret_cast<uint2*>(hi), reinterpret_cast<unsigned char*>(lo8), n4);
  return ok_or_cuda2();
}

int stream_decode_hl8(const void* hi, const void* lo8, float* out, long long n, cudaStream_t s) {
  if (n % 64 != 0) return DFIR_ERR_ARG;
  if (n == 0) return DFIR_OK;
  const long long n4 = n / 4;
  stream_decode_hl8_kernel<<<static_cast<unsigned>(std::min<long long>((n4 + 255) / 256, 148 * 8)), 256, 0, s>>>(
      reinterpret_cast<const uint2*>(hi), reinterpret_cast<const unsigned char*>(lo8), reinterpret_cast<float4*>(out), n4);
  return ok_or_cuda2();
}

int channel_scale(const float* x, const float* svec, const float* add, float alpha, float* out, int B, long long HW,
                  int C, cudaStream_t s) {
  if (C % 4 != 0) return DFIR_ERR_ARG;
  const long long n4 = static_cast<long long>(B) * HW * C / 4;
  if (n4 == 0) return DFIR_OK;
  channel_scale_kernel<<<static_cast<unsigned>(std::min<long long>((n4 + 255) / 256, 148 * 8)), 256, 0, s>>>(
      reinterpret_cast<const float4*>(x), svec, reinterpret_cast<const float4*>(add), alpha,
      reinterpret_cast<float4*>(out), HW * C / 4, C / 4, n4);
  return ok_or_cuda2();
}

int lam_forward(const float* stack, long long map_stride, float gamma, float* out, float* scratch, int N, int B, int HW,
                int C, cudaStream_t s) {
  if (N < 1 || N > kLamMaxN) return DFIR_ERR_ARG;
  if (B == 0) return DFIR_OK;
  const int nchunk = 32;
  float* partial = scratch;                                        // [B][nchunk][N*N]
  float* att = scratch + static_cast<size_t>(B) * nchunk * N * N;   // [B][N][N]
  lam_gram_kernel<<<dim3(nchunk, B), 256, 0, s>>>(stack, map_stride, partial, N, HW * C, nchunk);
  lam_attention_kernel<<<B, 64, 0, s>>>(partial, att, N, nchunk);
  lam_apply_kernel<<<dim3(64, B), 256, 0, s>>>(stack, map_stride, att, gamma, out, N, HW, C);
  return ok_or_cuda2();
}

int csam_forward(const float* x, const float* w27, float bias, float gamma, float* out, int B, int H, int W, int C,
                 cudaStream_t s) {
  const long long n = static_cast<long long>(B) * H * W * C;
  if (n == 0) return DFIR_OK;
  csam_kernel<<<static_cast<unsigned>(std::min<long long>((n + 255) / 256, 148 * 16)), 256, 0, s>>>(x, w27, bias, gamma,
                                                                                                    out, B, H, W, C);
  return ok_or_cuda2();
}

int soca_forward(const float* x, const float* mlp, int R, float* svec, float* scratch, int B, int H, int W, int C,
                 cudaStream_t s) {
  if (C != 64) return DFIR_ERR_ARG;
  if (B == 0) return DFIR_OK;
  // SOCA centre-crops to 1000 along any side that is >= 1000 (SAN_blocks.py:265-280; sides of exactly 1000 fall
  // into the reference's `else` branch, which crops a 1000-window starting at 0 — the same window)
  int y0 = 0, x0 = 0, h1 = H, w1 = W;
  if (H >= 1000) { y0 = (H - 1000) / 2; h1 = 1000; }
  if (W >= 1000) { x0 = (W - 1000) / 2; w1 = 1000; }
  if (!(H < 1000 && W < 1000) && !(H >= 1000 && W >= 1000)) {
    // mixed case: the reference crops only the long side (its first three branches)
    if (H < 1000) { y0 = 0; h1 = H; }
    if (W < 1000) { x0 = 0; w1 = W; }
  }
  const int nchunk = 32;
  float* partial = scratch;                                            // [B][nchunk][4096+64]
  float* cov = scratch + static_cast<size_t>(B) * nchunk * (4096 + 64); // [B][4096]
  covpool_partial_kernel<<<dim3(nchunk, B), 256, 0, s>>>(x, partial, H, W, y0, x0, h1, w1, nchunk);
  covpool_final_kernel<<<dim3(B, 16), 256, 0, s>>>(partial, cov, h1 * w1, nchunk);
  static bool configured[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return DFIR_ERR_CUDA;
  if (dev < 0 || dev >= 64 || !configured[dev]) {
    if (cudaFuncSetAttribute(soca_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 4096 * 4) != cudaSuccess)
      return DFIR_ERR_CUDA;
    if (dev >= 0 && dev < 64) configured[dev] = true;
  }
  soca_kernel<<<B, 256, 4 * 4096 * 4, s>>>(cov, mlp, R, svec, 5);
  return ok_or_cuda2();
}

static void soca_window(int H, int W, int* y0, int* x0, int* h1, int* w1) {
  // SOCA centre-crops to 1000 along any side that is >= 1000 (SAN_blocks.py:265-280)
  *y0 = 0; *x0 = 0; *h1 = H; *w1 = W;
  if (H >= 1000) { *y0 = (H - 1000) / 2; *h1 = 1000; }
  if (W >= 1000) { *x0 = (W - 1000) / 2; *w1 = 1000; }
}

constexpr int kCovChunks = 32;

size_t covpool_scratch_floats(int B) { return static_cast<size_t>(B) * kCovChunks * (4096 + 64); }

int covpool_forward(const float* x, float* cov, float* scratch, int B, int H, int W, int C, int crop1000, cudaStream_t s) {
  if (C != 64) return DFIR_ERR_ARG;
  if (B == 0) return DFIR_OK;
  int y0 = 0, x0 = 0, h1 = H, w1 = W;
  if (crop1000) soca_window(H, W, &y0, &x0, &h1, &w1);
  covpool_partial_kernel<<<dim3(kCovChunks, B), 256, 0, s>>>(x, scratch, H, W, y0, x0, h1, w1, kCovChunks);
  covpool_final_kernel<<<dim3(B, 16), 256, 0, s>>>(scratch, cov, h1 * w1, kCovChunks);
  return ok_or_cuda2();
}

int covpool_backward(const float* x, const float* grad_cov, float* grad_x, float* scratch, int B, int H, int W, int C,
                     int crop1000, cudaStream_t s) {
  if (C != 64) return DFIR_ERR_ARG;
  if (B == 0) return DFIR_OK;
  int y0 = 0, x0 = 0, h1 = H, w1 = W;
  if (crop1000) soca_window(H, W, &y0, &x0, &h1, &w1);
  if (h1 != H || w1 != W) {  // pixels outside the window receive no gradient
    if (cudaMemsetAsync(grad_x, 0, static_cast<size_t>(B) * H * W * 64 * sizeof(float), s) != cudaSuccess) return DFIR_ERR_CUDA;
  }
  channel_sum_partial_kernel<<<dim3(kCovChunks, B), 256, 0, s>>>(x, scratch, H, W, y0, x0, h1, w1, kCovChunks);
  const long long M = static_cast<long long>(h1) * w1;
  const int nchunk = static_cast<int>(std::max<long long>(1, std::min<long long>((M + 255) / 256, 148 * 4 / std::max(1, std::min(B, 8)) + 1)));
  covpool_bwd_kernel<<<dim3(nchunk, B), 256, 0, s>>>(x, grad_cov, scratch, grad_x, H, W, y0, x0, h1, w1, nchunk, kCovChunks);
  return ok_or_cuda2();
}

size_t sqrtm_scratch_floats(int B, int iters) { return static_cast<size_t>(B) * 2 * std::max(1, iters - 1) * 4096; }

static int sqrtm_configure(const void* fn, int bytes) {
  return cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) == cudaSuccess ? DFIR_OK : DFIR_ERR_CUDA;
}

int sqrtm_forward(const float* cov, float* out, int B, int C, int iters, cudaStream_t s) {
  if (C != 64 || iters < 2 || iters > 16) return DFIR_ERR_ARG;
  if (B == 0) return DFIR_OK;
  if (sqrtm_configure(reinterpret_cast<const void*>(sqrtm_fwd_kernel), 6 * 4096 * 4) != DFIR_OK) return DFIR_ERR_CUDA;
  sqrtm_fwd_kernel<<<B, 256, 6 * 4096 * 4, s>>>(cov, out, iters);
  return ok_or_cuda2();
}

int sqrtm_backward(const float* cov, const float* grad_out, float* grad_in, float* scratch, int B, int C, int iters,
                   cudaStream_t s) {
  if (C != 64 || iters < 2 || iters > 16) return DFIR_ERR_ARG;
  if (B == 0) return DFIR_OK;
  if (sqrtm_configure(reinterpret_cast<const void*>(sqrtm_bwd_kernel), 10 * 4096 * 4) != DFIR_OK) return DFIR_ERR_CUDA;
  sqrtm_bwd_kernel<<<B, 256, 10 * 4096 * 4, s>>>(cov, grad_out, grad_in, scratch, iters);
  return ok_or_cuda2();
}

size_t soca_scratch_floats(int B) { return static_cast<size_t>(B) * 32 * (4096 + 64) + static_cast<size_t>(B) * 4096; }
size_t lam_scratch_floats(int B, int N) { return static_cast<size_t>(B) * 32 * N * N + static_cast<size_t>(B) * N * N; }

int nonlocal_forward(const float* x, const float* wq, const float* bq, const float* wW, const float* bW, float* out,
                     float* scratch, int B, int H, int W, int C, cudaStream_t s) {
  if (C != 64) return DFIR_ERR_ARG;
  if (B == 0 || H < 2 || W < 2) return B == 0 ? DFIR_OK : DFIR_ERR_ARG;
  const long long npix = static_cast<long long>(B) * H * W;
  const int nk_max = ((H - H / 2) / 2) * ((W - W / 2) / 2) + 1;
  float* proj = scratch;                       // [npix][24]
  float* keys = scratch + npix * 24;           // [B][4][nk_max][16]
  nl_project_kernel<<<static_cast<unsigned>(std::min<long long>((npix + 255) / 256, 148 * 8)), 256, 0, s>>>(x, wq, bq, proj,
                                                                                                            npix);
  nl_pool_kernel<<<dim3(std::max(1, (nk_max + 127) / 128), 4, B), 128, 0, s>>>(proj, keys, B, H, W, nk_max);
  const int nq_max = (H - H / 2) * (W - W / 2);
  nl_attend_kernel<<<dim3((nq_max + 127) / 128, 4, B), 128, 0, s>>>(x, proj, keys, wW, bW, out, H, W, nk_max);
  return ok_or_cuda2();
}

// theta / phi / g projections and the pooled keys only (the backward pass rebuilds them instead of stashing them)
int nonlocal_project_pool(const float* x, const float* wq, const float* bq, float* proj, float* keys, int B, int H, int W,
                          cudaStream_t s) {
  if (B == 0) return DFIR_OK;
  if (H < 2 || W < 2) return DFIR_ERR_ARG;
  const long long npix = static_cast<long long>(B) * H * W;
  const int nk_max = ((H - H / 2) / 2) * ((W - W / 2) / 2) + 1;
  nl_project_kernel<<<static_cast<unsigned>(std::min<long long>((npix + 255) / 256, 148 * 8)), 256, 0, s>>>(x, wq, bq, proj,
                                                                                                            npix);
  nl_pool_kernel<<<dim3(std::max(1, (nk_max + 127) / 128), 4, B), 128, 0, s>>>(proj, keys, B, H, W, nk_max);
  return ok_or_cuda2();
}

size_t nonlocal_scratch_floats(int B, int H, int W) {
  const size_t npix = static_cast<size_t>(B) * H * W;
  const size_t nk_max = static_cast<size_t>((H - H / 2) / 2) * ((W - W / 2) / 2) + 1;
  return npix * 24 + static_cast<size_t>(B) * 4 * nk_max * 16;
}

}  // namespace dfir
